"""GPU-backed mirrors of the reference's batched entry points.

``ChunkPreprocessor`` has the constructor and ``run(chr_name, start, end)`` /
``process_items(items)`` contract of the reference class
(sai/preprocessors/chunk_preprocessor.py:38-147) and returns item dicts with
the schema of ``FeaturePreprocessor.run``
(sai/preprocessors/feature_preprocessor.py:119-191), in the reference's order
(population product outermost, windows innermost,
sai/generators/window_generator.py:164-167) -- but the per-window Python loop
is replaced by one call into the CUDA library per chunk.
"""

from __future__ import annotations

from itertools import combinations, product
from pathlib import Path
from typing import Any, Optional

import numpy as np

from . import _cabi
from .encode import PopData
from .scoring import HostEngine, dd_values, four_pop_values, make_job
from .windows import chunk_windows, split_genome


def _same_positions(members: list[PopData]) -> bool:
    first = members[0].POS
    return all(m.POS.shape == first.shape and np.array_equal(m.POS, first) for m in members[1:])


def _common_rows(members: list[PopData]) -> tuple[np.ndarray, list[np.ndarray]]:
    """Positions present in every member (window_generator.py:193-197) and each
    member's genotype rows at those positions (:217-231)."""
    if _same_positions(members):
        return members[0].POS, [m.GT for m in members]
    common = members[0].POS
    for m in members[1:]:
        common = np.intersect1d(common, m.POS)
    return common, [m.GT.compress(np.isin(m.POS, common), axis=0) for m in members]


def score_populations(
    chr_name: str,
    windows_by_tgt: dict[str, list[tuple[int, int]]],
    ref_data: dict[str, PopData],
    tgt_data: dict[str, PopData],
    src_data: dict[str, PopData],
    ploidy_config,
    stat_config,
    anc_allele_available: bool,
    engine: HostEngine,
    out_data: Optional[dict[str, PopData]] = None,
    num_src: Optional[int] = None,
) -> list[dict[str, Any]]:
    """Item dicts for every (ref, tgt, src-combination, outgroup) x window.

    When every population sits on the same positions and every target population has the same
    windows (always the case for ``ChunkPreprocessor.run(chr, start, end)`` on one VCF), ALL
    populations are packed once and the population product becomes up to ``SAI_MAX_JOBS`` jobs per
    genotype pass, instead of one pack + one transfer + one pass per combination."""
    stats = [s for s in stat_config.root.keys() if s in ("U", "Q")]
    four = [s for s in stat_config.root.keys() if s in ("Danc", "Dplus", "df", "fd") and stat_config.root[s] is True]
    dd = any(s == "DD" and stat_config.root[s] is True for s in stat_config.root.keys())
    num_src = len(src_data) if num_src is None else num_src
    src_combos = list(combinations(src_data.keys(), num_src))
    outs = list(out_data.keys()) if out_data else [None]
    src_ploidies = ploidy_config.get_ploidy("src")
    combos = list(product(ref_data, tgt_data, src_combos, outs))
    items: list[dict[str, Any]] = []
    if not combos:
        return items

    def job_for(ref_pop, tgt_pop, ref_idx, tgt_idx, src_idx):
        specs = {}
        for s in stats:
            prm = stat_config.get_parameters(s)
            spec = {"w": prm["ref"][ref_pop], "y_list": list(prm["src"].values())}
            spec["x" if s == "U" else "quantile"] = prm["tgt"][tgt_pop]
            specs[s] = spec
        return make_job(ref_idx, tgt_idx, src_idx, anc_allele_available, specs.get("U"), specs.get("Q"))

    def src_ploidy_list(n_src):
        # positional zip of src genotypes / y_list / ploidies, exactly as
        # feature_preprocessor.py:160,168-170 and stat_utils.py:116-119 do
        pl = list(src_ploidies[:n_src])
        if len(pl) != n_src:
            raise ValueError("The length of src_gts_list and src ploidies must match.")
        return pl

    stat_order = [s for s in stat_config.root.keys() if s in stats or s in four or (dd and s == "DD")]
    empty_cdd = np.array([])

    def emit(combo, windows, pos, res, j, four_vals, dd_vals):
        """One item dict per window (feature_preprocessor.py:119-191).  A chromosome-scale chunk has
        tens of thousands of windows, so the per-window work is kept to plain Python objects: the
        result columns are turned into lists once and the candidate lists are views into one
        converted array."""
        ref_pop, tgt_pop, src_comb, out_pop = combo
        n_src = len(src_comb)
        out_name = "NA" if out_pop is None else out_pop
        pos_dtype = np.asarray(pos).dtype
        if res is not None:
            nsnps_l = res.nsnps[j].tolist()
        else:
            nsnps_l = [int(np.count_nonzero((pos >= s_) & (pos <= e_))) for s_, e_ in windows]
        want_u, want_q = "U" in stats, "Q" in stats
        if want_u:
            u_l, us_l = res.u[j].tolist(), res.u_start[j].tolist()
            u_cand = res.u_cand[j].astype(pos_dtype, copy=False)
        if want_q:
            q_col, qnan_l = res.q[j], np.isnan(res.q[j]).tolist()  # q_col[i] is a numpy.float64, like np.nanquantile's
            qs_l, qc_l = res.q_start[j].tolist(), res.q_cnt[j].tolist()
            q_cand = res.q_cand[j].astype(pos_dtype, copy=False)
        nan_list = [np.nan for _ in range(n_src)] if n_src > 1 else np.nan
        for i, (start, end) in enumerate(windows):
            nsnps = nsnps_l[i]
            cdd = {}
            item = {
                "chr_name": chr_name,
                "start": start,
                "end": end,
                "ref_pop": ref_pop,
                "tgt_pop": tgt_pop,
                "src_pop_list": src_comb,
                "out_pop": out_name,
                "nsnps": nsnps,
                "cdd_pos": cdd,
            }
            for s in stat_order:
                if s == "U":
                    if nsnps == 0:  # empty window: feature_preprocessor.py:131-144
                        item[s] = np.nan
                        cdd[s] = empty_cdd
                    else:
                        n = u_l[i]
                        item[s] = n
                        cdd[s] = u_cand[us_l[i] : us_l[i] + n]
                elif s == "Q":
                    if nsnps == 0 or qnan_l[i]:  # q_statistic.py:96-98
                        item[s] = np.nan
                        cdd[s] = empty_cdd
                    else:
                        item[s] = q_col[i]
                        cdd[s] = q_cand[qs_l[i] : qs_l[i] + qc_l[i]]
                elif nsnps == 0:  # feature_preprocessor.py:137-141
                    item[s] = list(nan_list) if n_src > 1 else np.nan
                elif s == "DD":
                    item[s] = [dd_vals[k][i] for k in range(n_src)]
                else:
                    item[s] = [four_vals[s][k][i] for k in range(n_src)]
            items.append(item)

    def extras(pg, ref_idx, tgt_idx, src_idx, out_idx, n_ref, n_tgt, n_srcs):
        four_vals = dd_vals = None
        if four:
            four_vals = four_pop_values(engine.pattern_sums(pg, ref_idx, tgt_idx, out_idx, src_idx))
        if dd:
            ref_sum, tgt_sum = engine.dd_sums(pg, ref_idx, tgt_idx, src_idx)
            dd_vals = dd_values(ref_sum, tgt_sum, n_ref, n_tgt, n_srcs)
        return four_vals, dd_vals

    # ---- all populations in one packed matrix, the product as fused jobs ------------------
    use_out = bool(out_data) and bool(four)
    everyone = list(ref_data.values()) + list(tgt_data.values()) + list(src_data.values())
    if use_out:
        everyone += list(out_data.values())
    win_lists = list(windows_by_tgt.values())
    fused = (
        len(combos) > 1
        and len(src_combos) == 1
        and len(everyone) <= _cabi.MAX_POPS
        and _same_positions(everyone)
        and all(w == win_lists[0] for w in win_lists[1:])
        and (stats or four or dd)
    )
    if fused:
        index, mats, ploidy = {}, [], []
        for group, data in (("ref", ref_data), ("tgt", tgt_data)):
            for name, d in data.items():
                index[(group, name)] = len(mats)
                mats.append(d.GT)
                ploidy.append(ploidy_config.get_ploidy(group, name))
        src_comb = src_combos[0]
        for name, pl in zip(src_comb, src_ploidy_list(len(src_comb))):
            index[("src", name)] = len(mats)
            mats.append(src_data[name].GT)
            ploidy.append(pl)
        if use_out:
            for name, d in out_data.items():
                index[("out", name)] = len(mats)
                mats.append(d.GT)
                ploidy.append(ploidy_config.get_ploidy("outgroup", name))
        pos = everyone[0].POS
        windows = win_lists[0]
        src_idx = [index[("src", s)] for s in src_comb]
        pg = None
        for b0 in range(0, len(combos), _cabi.MAX_JOBS):
            batch = combos[b0 : b0 + _cabi.MAX_JOBS]
            jobs = [job_for(r, t, index[("ref", r)], index[("tgt", t)], src_idx) for r, t, _, _ in batch]
            if pg is None:  # int8 matrices -> pack | copy | genotype pass, pipelined; ONE upload for all batches
                res, pg = engine.score_matrices(mats, ploidy, pos, windows, jobs, keep_negatives=dd)
            else:
                res = engine.score_resident(jobs)
            for j, combo in enumerate(batch):
                r, t, _, o = combo
                four_vals, dd_vals = extras(
                    pg, index[("ref", r)], index[("tgt", t)], src_idx, index[("out", o)] if (use_out and o is not None) else -1,
                    ref_data[r].GT.shape[1], tgt_data[t].GT.shape[1], [src_data[s].GT.shape[1] for s in src_comb])
                emit(combo, windows, pos, res, j, four_vals, dd_vals)
        return items

    # ---- general case: one packed matrix per combination ----------------------------------
    for combo in combos:
        ref_pop, tgt_pop, src_comb, out_pop = combo
        windows = windows_by_tgt[tgt_pop]
        members = [ref_data[ref_pop], tgt_data[tgt_pop]] + [src_data[s] for s in src_comb]
        if out_pop is not None:
            members.append(out_data[out_pop])
        pos, rows = _common_rows(members)
        n_src = len(src_comb)
        ploidy = [ploidy_config.get_ploidy("ref", ref_pop), ploidy_config.get_ploidy("tgt", tgt_pop)]
        ploidy += src_ploidy_list(n_src)
        src_idx = list(range(2, 2 + n_src))
        job = job_for(ref_pop, tgt_pop, 0, 1, src_idx)
        n_pack = 2 + n_src
        if out_pop is not None and four:
            ploidy.append(ploidy_config.get_ploidy("outgroup", out_pop))
            n_pack += 1
        res = pg = None
        if stats or four or dd:
            res, pg = engine.score_matrices(rows[:n_pack], ploidy[:n_pack], pos, windows, [job], keep_negatives=dd)
        four_vals, dd_vals = extras(pg, 0, 1, src_idx, 2 + n_src if (out_pop is not None and four) else -1,
                                    rows[0].shape[1], rows[1].shape[1], [rows[2 + k].shape[1] for k in range(n_src)])
        emit(combo, windows, pos, res, 0, four_vals, dd_vals)
    return items


def write_items(output_file: str, items: list[dict[str, Any]], stat_config) -> None:
    """Appends score rows and ``.U.log`` / ``.Q.log`` rows with the reference's
    text layout (feature_preprocessor.py:193-258); lines are assembled in memory and written
    with one call per file."""
    names = [s for s in stat_config.root.keys() if s in ("U", "Q") or stat_config.root[s] is True]
    rows = []
    for it in items:
        parts = []
        n_src = len(it["src_pop_list"])
        for s in names:
            v = it.get(s)
            if isinstance(v, list) and len(v) == n_src:
                parts.extend("" if x is None else str(x) for x in v)
            else:
                if isinstance(v, list):
                    v = v[0] if v else ""
                parts.append("" if v is None else str(v))
        rows.append(
            f"{it['chr_name']}\t{it['start']}\t{it['end']}\t{it['ref_pop']}\t{it['tgt_pop']}\t"
            f"{','.join(it['src_pop_list'])}\t{it['out_pop']}\t{it['nsnps']}\t" + "\t".join(parts) + "\n"
        )
    with open(output_file, "a") as f:
        f.write("".join(rows))
    for key in ("U", "Q"):
        if key not in stat_config.root:
            continue
        rows = []
        for it in items:
            c = it["cdd_pos"][key]
            if c.size == 0:
                txt = "NA"
            else:
                prefix = f"{it['chr_name']}:"
                txt = ",".join([prefix + p for p in map(str, c.tolist())])
            rows.append(f"{it['chr_name']}\t{it['start']}\t{it['end']}\t{txt}\n")
        with open(Path(output_file).with_suffix(f".{key}.log"), "a") as f:
            f.write("".join(rows))


class ChunkPreprocessor:
    """Drop-in for ``sai.preprocessors.ChunkPreprocessor`` on the U/Q path."""

    def __init__(
        self,
        vcf_file: str,
        ref_ind_file: str,
        tgt_ind_file: str,
        src_ind_file: str,
        out_ind_file: Optional[str],
        win_len: int,
        win_step: int,
        output_file: str,
        ploidy_config,
        stat_config,
        anc_allele_file: Optional[str] = None,
        num_src: int = 1,
        device: int = 0,
    ):
        from .vcf import parse_ind_file

        self.vcf_file = vcf_file
        self.ref_ind_file = ref_ind_file
        self.tgt_ind_file = tgt_ind_file
        self.src_ind_file = src_ind_file
        self.out_ind_file = out_ind_file
        self.win_len = win_len
        self.win_step = win_step
        self.output_file = output_file
        self.ploidy_config = ploidy_config
        self.stat_config = stat_config
        self.anc_allele_file = anc_allele_file
        self.num_src = len(parse_ind_file(src_ind_file).keys())
        self.anc_allele_available = anc_allele_file is not None
        self.engine = HostEngine(device)  # picklable; CUDA context created in run()

    def run(self, chr_name: str, start: int, end: int) -> list[dict[str, Any]]:
        from .vcf import read_data

        if self.win_len <= 0:
            raise ValueError("`win_len` must be greater than 0.")
        if self.win_step < 0:
            raise ValueError("`win_step` must be non-negative.")
        groups = read_data(
            vcf_file=self.vcf_file,
            chr_name=chr_name,
            ploidy_config=self.ploidy_config,
            ref_ind_file=self.ref_ind_file,
            tgt_ind_file=self.tgt_ind_file,
            src_ind_file=self.src_ind_file,
            out_ind_file=self.out_ind_file,
            anc_allele_file=self.anc_allele_file,
            start=start,
            end=end,
        )
        ref_data, ref_samples = groups["ref"]
        tgt_data, tgt_samples = groups["tgt"]
        src_data, src_samples = groups["src"]
        out_data, out_samples = groups["outgroup"]
        if start is None and end is None:
            windows = {t: split_genome(tgt_data[t].POS, self.win_len, self.win_step) for t in tgt_samples}
        else:
            wins = chunk_windows(start, end, self.win_len, self.win_step)
            windows = {t: wins for t in tgt_samples}
        if ref_data is None or tgt_data is None or src_data is None:
            return self._empty_items(chr_name, windows, ref_samples, tgt_samples, src_samples, out_samples)
        return score_populations(
            chr_name, windows, ref_data, tgt_data, src_data, self.ploidy_config, self.stat_config,
            self.anc_allele_available, self.engine, out_data=out_data, num_src=self.num_src,
        )

    def _empty_items(self, chr_name, windows, ref_samples, tgt_samples, src_samples, out_samples=None):
        # no data in the region: window_generator.py:249-289 (product over the outgroup populations
        # too, :269-273) + feature_preprocessor.py:131-144
        stats = [s for s in self.stat_config.root.keys() if s in ("U", "Q")]
        four = [s for s in self.stat_config.root.keys() if s in ("Danc", "Dplus", "df", "fd", "DD")]
        items = []
        for ref_pop, tgt_pop, src_comb, out_pop in product(
            ref_samples, tgt_samples, list(combinations(src_samples.keys(), self.num_src)), out_samples or [None]
        ):
            for start, end in windows[tgt_pop]:
                it = {
                    "chr_name": chr_name, "start": start, "end": end, "ref_pop": ref_pop, "tgt_pop": tgt_pop,
                    "src_pop_list": src_comb, "out_pop": "NA" if out_pop is None else out_pop, "nsnps": 0,
                    "cdd_pos": {},
                }
                for s in four:
                    it[s] = [np.nan for _ in src_comb] if len(src_comb) > 1 else np.nan
                for s in stats:
                    it[s] = np.nan
                    it["cdd_pos"][s] = np.array([])
                items.append(it)
        return items

    def process_items(self, items: list[dict[str, Any]]) -> None:
        write_items(self.output_file, items, self.stat_config)
