"""CPU oracle for the U / Q sliding-window scoring path of xin-huang/sai.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sai_b200/`` imports this module; it
is used by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the *checker*
and the *timed CPU baseline*, never as the product path.

This is a numpy restatement (not a copy) of the reference's algorithm.  Each
function cites the reference file:line (relative to the upstream repository
root) whose behaviour it follows.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this module
against (i) the known-answer vectors in the reference's own unit tests
(``tests/stats/test_u_statistic.py``, ``tests/stats/test_q_statistic.py``,
``tests/stats/test_stat_utils.py``, ``tests/utils/test_utils.py``,
``tests/generators/test_chunk_generator.py``), (ii) the pipeline goldens on
``tests/data/example.vcf`` (U == 3, Q == 0.9) and (iii) fixtures produced by
importing the reference's own ``sai.stats`` / ``WindowGenerator`` /
``FeaturePreprocessor`` in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz|json``).

Third-party arithmetic on the path: ``numpy.nanquantile`` (default
``method="linear"``), numpy pinned to 1.26.4 by the reference
(``pyproject.toml:23``); the formula (virtual index ``(n-1)*q``, ``_lerp`` with
the ``t >= 0.5`` branch) is unchanged between 1.22 and the 2.3.x installed
here.  ``quantile_linear`` below restates it explicitly and is checked against
``numpy.nanquantile`` itself.
"""

from __future__ import annotations

import math
from itertools import combinations, product
from typing import Any, Iterator, Optional, Sequence

import numpy as np

_OPS = ("=", "<", ">", "<=", ">=")


# --------------------------------------------------------------------------
# A1  per-site frequency           (sai/stats/stat_utils.py:26-52)
# --------------------------------------------------------------------------
def site_frequency(gts: np.ndarray, ploidy: int = 1) -> np.ndarray:
    """Alt-allele frequency per site of one population.

    ``gts`` is ``(sites, individuals)`` of per-individual allele sums; any
    negative entry is a missing call.  ``freq = sum(called values) /
    (n_called * ploidy)`` in float64, NaN where nobody is called
    (stat_utils.py:45-52).  Non positive-int ploidy raises (stat_utils.py:42-43).
    """
    if not isinstance(ploidy, int) or ploidy <= 0:
        raise ValueError("ploidy must be a positive integer.")
    gts = np.asarray(gts)
    present = gts >= 0
    n_called = present.sum(axis=1)
    numer = (gts * present).sum(axis=1, dtype=float)
    denom = n_called * ploidy
    freq = np.full(gts.shape[0], np.nan, dtype=float)
    ok = denom > 0
    freq[ok] = numer[ok] / denom[ok]
    return freq


def site_counts(gts: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Integer ``(num, called)`` per site -- the quantities the CUDA site-count
    kernel emits; ``site_frequency == num / (called * ploidy)``."""
    gts = np.asarray(gts)
    present = gts >= 0
    return (gts * present).sum(axis=1).astype(np.int64), present.sum(axis=1).astype(
        np.int64
    )


def _compare(op: str, freq: np.ndarray, y: float) -> np.ndarray:
    if op == "=":
        return freq == y
    if op == "<":
        return freq < y
    if op == ">":
        return freq > y
    if op == "<=":
        return freq <= y
    return freq >= y


# --------------------------------------------------------------------------
# A2  site condition               (sai/stats/stat_utils.py:55-168)
# --------------------------------------------------------------------------
def matching_loci(
    ref_gts,
    tgt_gts,
    src_gts_list,
    w: float,
    y_list: Sequence[tuple[str, float]],
    ploidy: Sequence[int],
    anc_allele_available: bool,
):
    """Returns ``(ref_freq, tgt_freq, condition)`` like the reference's
    ``compute_matching_loci``.

    Validation order and messages follow stat_utils.py:99-111.  ``valid`` is
    "every frequency finite and inside [0, 1]" (:121-130).  Every src must
    satisfy its comparator against ``y`` (:133-144); without ancestral alleles
    the comparators are also tried against the Python float ``1 - y`` and
    sites matching *that* have ref/tgt frequencies replaced by ``1 - freq``
    (:146-160).  ``condition = valid & match & (ref_freq < w)`` (:166).
    """
    if not (0 <= w <= 1):
        raise ValueError("Parameters w must be within the range [0, 1].")
    for op, y in y_list:
        if not (0 <= y <= 1):
            raise ValueError(f"Invalid value in y_list: {y}. within the range [0, 1].")
        if op not in _OPS:
            raise ValueError(
                f"Invalid operator in y_list: {op}. Must be '=', '<', '>', '<=', or '>='."
            )
    if len(src_gts_list) != len(y_list):
        raise ValueError("The length of src_gts_list and y_list must match.")

    ref_freq = site_frequency(ref_gts, ploidy[0])
    tgt_freq = site_frequency(tgt_gts, ploidy[1])
    src_freqs = [site_frequency(g, p) for g, p in zip(src_gts_list, ploidy[2:])]

    def in_unit(f):
        return np.isfinite(f) & (f >= 0) & (f <= 1)

    valid = in_unit(ref_freq) & in_unit(tgt_freq)
    for f in src_freqs:
        valid &= in_unit(f)

    with np.errstate(invalid="ignore"):
        match_y = np.ones(ref_freq.shape[0], dtype=bool)
        for f, (op, y) in zip(src_freqs, y_list):
            match_y &= _compare(op, f, y)
        if anc_allele_available:
            match = match_y
        else:
            match_flip = np.ones(ref_freq.shape[0], dtype=bool)
            for f, (op, y) in zip(src_freqs, y_list):
                match_flip &= _compare(op, f, 1 - y)
            match = match_y | match_flip
            flip = match_flip & valid
            ref_freq[flip] = 1 - ref_freq[flip]
            tgt_freq[flip] = 1 - tgt_freq[flip]
        condition = valid & match & (ref_freq < w)
    return ref_freq, tgt_freq, condition


# --------------------------------------------------------------------------
# A3  U                            (sai/stats/u_statistic.py:70-99)
# --------------------------------------------------------------------------
def u_statistic(
    ref_gts, tgt_gts, src_gts_list, ref_ploidy, tgt_ploidy, src_ploidy_list, **kw
) -> dict[str, Any]:
    need = ["pos", "w", "x", "y_list", "anc_allele_available"]
    absent = [k for k in need if k not in kw]
    if absent:
        raise ValueError(f"Missing required argument(s): {', '.join(absent)}")
    _, tgt_freq, cond = matching_loci(
        ref_gts,
        tgt_gts,
        src_gts_list,
        kw["w"],
        kw["y_list"],
        [ref_ploidy, tgt_ploidy] + list(src_ploidy_list),
        kw["anc_allele_available"],
    )
    with np.errstate(invalid="ignore"):
        cond = cond & (tgt_freq > kw["x"])
    hits = np.flatnonzero(cond)
    return {"name": "U", "value": hits.size, "cdd_pos": np.asarray(kw["pos"])[hits]}


# --------------------------------------------------------------------------
# numpy 'linear' quantile, restated (numpy/lib/_function_base_impl.py
# _get_indexes / _get_gamma / _lerp; pinned numpy==1.26.4, same formula)
# --------------------------------------------------------------------------
def quantile_linear(values: np.ndarray, q: float) -> float:
    """Explicit form of ``numpy.quantile(values, q)`` for a non-empty, NaN-free
    float64 vector: ``vi = (n-1)*q``; above ``n-1`` -> max; else
    ``a=v[floor(vi)], b=v[floor(vi)+1], g=vi-floor(vi)`` and
    ``a + (b-a)*g`` if ``g < 0.5`` else ``b - (b-a)*(1-g)``, each operation
    rounded separately (no FMA)."""
    v = np.sort(np.asarray(values, dtype=np.float64))
    n = v.size
    vi = np.float64(n - 1) * np.float64(q)
    if vi >= n - 1:
        return float(v[-1])
    lo = math.floor(vi)
    g = np.float64(vi - lo)
    a, b = v[lo], v[lo + 1]
    d = np.float64(b - a)
    if g >= 0.5:
        return float(b - np.float64(d * np.float64(1 - g)))
    return float(a + np.float64(d * g))


# --------------------------------------------------------------------------
# A4  Q                            (sai/stats/q_statistic.py:70-104)
# --------------------------------------------------------------------------
def q_statistic(
    ref_gts, tgt_gts, src_gts_list, ref_ploidy, tgt_ploidy, src_ploidy_list, **kw
) -> dict[str, Any]:
    need = ["pos", "w", "y_list", "anc_allele_available", "quantile"]
    absent = [k for k in need if k not in kw]
    if absent:
        raise ValueError(f"Missing required argument(s): {', '.join(absent)}")
    _, tgt_freq, cond = matching_loci(
        ref_gts,
        tgt_gts,
        src_gts_list,
        kw["w"],
        kw["y_list"],
        [ref_ploidy, tgt_ploidy] + list(src_ploidy_list),
        kw["anc_allele_available"],
    )
    kept = tgt_freq[cond]
    kept_pos = np.asarray(kw["pos"])[cond]
    if kept.size == 0:
        return {"name": "Q", "value": np.nan, "cdd_pos": np.array([])}
    thr = np.nanquantile(kept, kw["quantile"])
    return {"name": "Q", "value": thr, "cdd_pos": kept_pos[kept >= thr]}


# --------------------------------------------------------------------------
# N3  Danc / Dplus / df / fd      (sai/stats/stat_utils.py:171-272 and the four
#     *_statistic.py classes)
# --------------------------------------------------------------------------
def pattern_sum(freqs, pattern: str) -> float:
    """Sum over sites of the product of ``f`` ('b') or ``1 - f`` ('a') of
    (ref, tgt, src, out), multiplied in that order starting from ones
    (stat_utils.py:258-272)."""
    if len(pattern) != 4:
        raise ValueError("Pattern must be a four-character string.")
    prod = np.ones_like(freqs[0])
    for f, c in zip(freqs, pattern.lower()):
        if c == "a":
            prod = prod * (1 - f)
        elif c == "b":
            prod = prod * f
        else:
            raise ValueError(f"Invalid character '{c}' in pattern. Only 'a' and 'b' allowed.")
    return float(np.sum(prod))


def pairwise_sum_model(a) -> float:
    """``float(np.sum(a))`` for a contiguous float64 vector, spelled out: numpy's pairwise
    summation (numpy/core/src/umath/loops_utils.h.src ``DOUBLE_pairwise_sum``; block size 128,
    eight accumulators; unchanged between the reference's numpy 1.26 and 2.x).  ``pattern_sum``
    calls ``np.sum`` itself; this restatement documents the order the CUDA kernel reproduces and
    is pinned against numpy in tests/test_oracle_golden.py."""
    a = [float(x) for x in a]

    def pw(lo: int, n: int) -> float:
        if n < 8:
            res = 0.0
            for i in range(lo, lo + n):
                res = res + a[i]
            return res
        if n <= 128:
            r = a[lo : lo + 8]
            i = 8
            while i < n - (n % 8):
                for j in range(8):
                    r[j] = r[j] + a[lo + i + j]
                i += 8
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
            while i < n:
                res = res + a[lo + i]
                i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return pw(lo, n2) + pw(lo + n2, n - n2)

    return pw(0, len(a))


def four_pop_statistics(ref_gts, tgt_gts, src_gts_list, ref_ploidy, tgt_ploidy, src_ploidy_list,
                        out_gts=None, out_ploidy=None) -> dict[str, list[float]]:
    """``{"Danc": [...], "Dplus": [...], "df": [...], "fd": [...]}``, one value per
    source population (danc_statistic.py:62-83, dplus_statistic.py:63-86,
    df_statistic.py:62-84, fd_statistic.py:63-89).  The outgroup frequency is
    0 when there is no outgroup (stat_utils.py:212-213)."""
    out = {"Danc": [], "Dplus": [], "df": [], "fd": []}
    fr = site_frequency(ref_gts, ref_ploidy)
    ft = site_frequency(tgt_gts, tgt_ploidy)
    fo = np.zeros_like(fr) if out_gts is None else site_frequency(out_gts, out_ploidy)
    ratio = lambda n, d: n / d if d != 0 else np.nan
    for src, sp in zip(src_gts_list, src_ploidy_list):
        fs = site_frequency(src, sp)
        four = (fr, ft, fs, fo)
        abba, baba = pattern_sum(four, "abba"), pattern_sum(four, "baba")
        baaa, abaa, bbaa = pattern_sum(four, "baaa"), pattern_sum(four, "abaa"), pattern_sum(four, "bbaa")
        dnr = np.maximum(ft, fs)
        abba_d, baba_d = pattern_sum((fr, dnr, dnr, fo), "abba"), pattern_sum((fr, dnr, dnr, fo), "baba")
        out["Danc"].append(ratio(baaa - abaa, baaa + abaa))
        out["Dplus"].append(ratio(abba - baba + baaa - abaa, abba + baba + baaa + abaa))
        out["df"].append(ratio(abba - baba, abba + baba + 2 * bbaa))
        out["fd"].append(ratio(abba - baba, abba_d - baba_d))
    return out


# --------------------------------------------------------------------------
# N4  DD                           (sai/stats/dd_statistic.py:62-77)
# --------------------------------------------------------------------------
def cityblock_sums(a_gts: np.ndarray, b_gts: np.ndarray) -> np.ndarray:
    """``out[i, j] = sum_sites |a[site, i] - b[site, j]|`` as float64 -- what
    ``scipy.spatial.distance.cdist(a.T, b.T, metric="cityblock")`` returns
    (dd_statistic.py:70-71); the raw values enter, negative missing codes
    included.  Integer-valued, hence exact in float64 in any summation order."""
    a = np.asarray(a_gts, dtype=np.float64)
    b = np.asarray(b_gts, dtype=np.float64)
    out = np.zeros((a.shape[1], b.shape[1]), dtype=np.float64)
    for i in range(a.shape[1]):
        out[i] = np.abs(a[:, i : i + 1] - b).sum(axis=0)
    return out


def dd_statistic(ref_gts, tgt_gts, src_gts_list) -> list:
    """One DD value per source population: the mean over source individuals of
    (mean distance to the ref individuals - mean distance to the tgt individuals)
    (dd_statistic.py:66-79)."""
    values = []
    for src_gts in src_gts_list:
        mean_src_tgt = np.mean(cityblock_sums(src_gts, tgt_gts), axis=1)
        mean_src_ref = np.mean(cityblock_sums(src_gts, ref_gts), axis=1)
        values.append(np.mean(mean_src_ref - mean_src_tgt))
    return values


# --------------------------------------------------------------------------
# A5  window grid                  (sai/utils/utils.py:558-612)
# --------------------------------------------------------------------------
def split_genome(pos, window_size: int, step_size: int, start: Optional[int] = None):
    if step_size <= 0 or window_size <= 0:
        raise ValueError("`step_size` and `window_size` must be positive integers.")
    if step_size > window_size:
        raise ValueError("`step_size` cannot be greater than `window_size`.")
    if len(pos) == 0:
        raise ValueError("`pos` array must not be empty.")
    first = (pos[0] + step_size) // step_size * step_size - window_size + 1
    first = max(first, 1 if start is None else start)
    out = []
    s = first
    while s <= pos[-1]:
        out.append((s, s + window_size - 1))
        s += step_size
    return out


# --------------------------------------------------------------------------
# sharding                         (sai/generators/chunk_generator.py:111-142)
# --------------------------------------------------------------------------
def split_windows_ranges(windows: list, num_chunks: int) -> list:
    base, extra = divmod(len(windows), num_chunks)
    out, at = [], 0
    for i in range(num_chunks):
        nxt = at + base + (1 if i < extra else 0)
        part = windows[at:nxt]
        if part:
            out.append((part[0][0], part[-1][1]))
        at = nxt
    return out


def chunk_windows(start: int, end: int, win_len: int, win_step: int):
    """Windows a chunk worker re-derives from its ``(start, end)`` range
    (sai/generators/window_generator.py:132-144)."""
    return split_genome([start, end - win_len + win_step], win_len, win_step, start=start)


# --------------------------------------------------------------------------
# A5  window extraction            (sai/generators/window_generator.py:150-247)
# --------------------------------------------------------------------------
class PopData:
    """Minimal stand-in for the reference's ``ChromosomeData`` (POS, GT)."""

    __slots__ = ("POS", "GT")

    def __init__(self, POS, GT):
        self.POS = np.asarray(POS)
        self.GT = np.asarray(GT)


def iter_windows(
    chr_name: str,
    ref_data: dict,
    tgt_data: dict,
    src_data: dict,
    windows_by_tgt: dict,
    num_src: int,
    ploidy_config,
    out_data: Optional[dict] = None,
) -> Iterator[dict[str, Any]]:
    """Yields the per-window dicts of the reference's ``_window_generator``:
    population product outermost, windows innermost; a window keeps the
    positions present (inclusive ``[start, end]``) in *every* population; an
    empty window yields ``pos=[]`` and ``None`` genotypes."""
    src_combos = list(combinations(src_data.keys(), num_src))
    outs = list(out_data.keys()) if out_data else [None]
    for ref_pop, tgt_pop, src_comb, out_pop in product(
        ref_data, tgt_data, src_combos, outs
    ):
        members = [ref_data[ref_pop], tgt_data[tgt_pop]] + [src_data[s] for s in src_comb]
        if out_pop is not None:
            members.append(out_data[out_pop])
        for start, end in windows_by_tgt[tgt_pop]:
            common = None
            for d in members:
                inside = d.POS[(d.POS >= start) & (d.POS <= end)]
                common = np.unique(inside) if common is None else np.intersect1d(common, inside)
            base = {
                "chr_name": chr_name,
                "ref_pop": ref_pop,
                "tgt_pop": tgt_pop,
                "src_pop_list": src_comb,
                "out_pop": out_pop,
                "start": start,
                "end": end,
                "ploidy_config": ploidy_config,
            }
            if common.size == 0:
                base.update(pos=[], ref_gts=None, tgt_gts=None, src_gts_list=None, out_gts=None)
                yield base
                continue
            picked = [d.GT.compress(np.isin(d.POS, common), axis=0) for d in members]
            base.update(
                pos=common,
                ref_gts=picked[0],
                tgt_gts=picked[1],
                src_gts_list=picked[2 : 2 + len(src_comb)],
                out_gts=picked[-1] if out_pop is not None else None,
            )
            yield base


# --------------------------------------------------------------------------
# A6  dispatch + item schema       (sai/preprocessors/feature_preprocessor.py:63-191)
# --------------------------------------------------------------------------
def window_item(win: dict, stat_config, anc_allele_available: bool) -> dict[str, Any]:
    """One output item for one window dict.  ``stat_config`` needs ``.root``
    (ordered mapping) and ``.get_parameters(name)``; ``win['ploidy_config']``
    needs ``.get_ploidy(group, pop=None)``.  U, Q and the four site-pattern
    statistics (Danc, Dplus, df, fd) and DD are covered."""
    item = {
        "chr_name": win["chr_name"],
        "start": win["start"],
        "end": win["end"],
        "ref_pop": win["ref_pop"],
        "tgt_pop": win["tgt_pop"],
        "src_pop_list": win["src_pop_list"],
        "out_pop": "NA" if win["out_pop"] is None else win["out_pop"],
        "nsnps": len(win["pos"]),
        "cdd_pos": {},
    }
    stats = [s for s in stat_config.root.keys() if s in ("U", "Q")]
    four = [s for s in stat_config.root.keys() if s in ("Danc", "Dplus", "df", "fd", "DD") and stat_config.root[s] is True]
    n_src = len(win["src_pop_list"])
    if win["ref_gts"] is None or win["tgt_gts"] is None or win["src_gts_list"] is None:
        for s in stat_config.root.keys():  # feature_preprocessor.py:137-144, in config order
            if s in four:
                item[s] = [np.nan] * n_src if n_src > 1 else np.nan
            elif s in stats:
                item[s] = np.nan
                item["cdd_pos"][s] = np.array([])
        return item
    pc = win["ploidy_config"]
    pops = dict(
        ref_gts=win["ref_gts"],
        tgt_gts=win["tgt_gts"],
        src_gts_list=win["src_gts_list"],
        ref_ploidy=pc.get_ploidy("ref", win["ref_pop"]),
        tgt_ploidy=pc.get_ploidy("tgt", win["tgt_pop"]),
        src_ploidy_list=pc.get_ploidy("src"),
    )
    four_vals = None
    if [s for s in four if s != "DD"]:
        four_vals = four_pop_statistics(**pops, out_gts=win["out_gts"],
                                        out_ploidy=pc.get_ploidy("outgroup", win["out_pop"]) if win["out_pop"] is not None else None)
    for s in stat_config.root.keys():
        if s == "DD" and s in four:
            item[s] = dd_statistic(win["ref_gts"], win["tgt_gts"], win["src_gts_list"])
            continue
        if s in four:
            item[s] = four_vals[s]
            continue
        if s not in stats:
            continue
        prm = stat_config.get_parameters(s)
        common = dict(
            pos=win["pos"],
            w=prm["ref"][win["ref_pop"]],
            y_list=list(prm["src"].values()),
            anc_allele_available=anc_allele_available,
        )
        if s == "U":
            res = u_statistic(**pops, x=prm["tgt"][win["tgt_pop"]], **common)
        else:
            res = q_statistic(**pops, quantile=prm["tgt"][win["tgt_pop"]], **common)
        item["cdd_pos"][s] = res["cdd_pos"]
        item[s] = res["value"]
    return item


def score_chunk(
    chr_name: str,
    start: int,
    end: int,
    win_len: int,
    win_step: int,
    ref_data: dict,
    tgt_data: dict,
    src_data: dict,
    ploidy_config,
    stat_config,
    anc_allele_available: bool,
    out_data: Optional[dict] = None,
) -> list[dict[str, Any]]:
    """What ``ChunkPreprocessor.run(chr_name, start, end)`` returns
    (sai/preprocessors/chunk_preprocessor.py:105-147) given in-memory
    population data already restricted to the chunk's region."""
    wins = chunk_windows(start, end, win_len, win_step)
    gen = iter_windows(
        chr_name,
        ref_data,
        tgt_data,
        src_data,
        {t: wins for t in tgt_data},
        num_src=len(src_data),
        ploidy_config=ploidy_config,
        out_data=out_data,
    )
    return [window_item(w, stat_config, anc_allele_available) for w in gen]


# --------------------------------------------------------------------------
# A7  text output                  (sai/preprocessors/feature_preprocessor.py:193-258,
#                                   headers sai/sai.py:111-144)
# --------------------------------------------------------------------------
def format_items(items: list[dict], stat_names: Sequence[str]):
    """Returns ``(score_lines, {"U": log_lines, "Q": log_lines})`` with the
    reference's exact text layout (``str(value)`` for every statistic)."""
    rows = []
    for it in items:
        parts = []
        for s in stat_names:
            v = it.get(s)
            if isinstance(v, list) and len(v) == len(it["src_pop_list"]):
                parts.extend("" if x is None else str(x) for x in v)  # one column per source
            else:
                if isinstance(v, list):
                    v = v[0] if v else ""
                parts.append("" if v is None else str(v))
        vals = "\t".join(parts)
        rows.append(
            f"{it['chr_name']}\t{it['start']}\t{it['end']}\t{it['ref_pop']}\t"
            f"{it['tgt_pop']}\t{','.join(it['src_pop_list'])}\t{it['out_pop']}\t"
            f"{it['nsnps']}\t{vals}\n"
        )
    logs = {}
    for key in ("U", "Q"):
        if key not in stat_names:
            continue
        lines = []
        for it in items:
            c = it["cdd_pos"][key]
            txt = "NA" if c.size == 0 else ",".join(f"{it['chr_name']}:{p}" for p in c)
            lines.append(f"{it['chr_name']}\t{it['start']}\t{it['end']}\t{txt}\n")
        logs[key] = lines
    return rows, logs


def score_header(stat_names: Sequence[str], src_pops: Sequence[str] = ()) -> str:
    """Header of the score file (sai/sai.py:111-131): U and Q one column each,
    the other statistics one column per source population when there are several."""
    cols = ["Chrom", "Start", "End", "Ref", "Tgt", "Src", "Outgroup", "N(Variants)"]
    for s in stat_names:
        if s in ("U", "Q") or len(src_pops) <= 1:
            cols.append(s)
        else:
            cols.extend(f"{s}.{sp}" for sp in src_pops)
    return "\t".join(cols) + "\n"


# --------------------------------------------------------------------------
# N1  outlier threshold            (sai/sai.py:192-214)
# --------------------------------------------------------------------------
def outlier_threshold(values: np.ndarray, q: float) -> Optional[float]:
    """Threshold ``sai outlier`` applies to one metric column: NaNs dropped;
    ``None`` when the column is empty or has one unique value (the reference
    warns and writes an empty table); else the linear quantile
    (pandas ``Series.quantile`` -> numpy linear)."""
    v = np.asarray(values, dtype=np.float64)
    v = v[~np.isnan(v)]
    if v.size == 0 or np.unique(v).size == 1:
        return None
    return float(np.quantile(v, q))


def outlier_mask(values: np.ndarray, q: float, column: str) -> np.ndarray:
    """Rows kept by ``sai outlier``: strict ``>`` for columns starting with
    "U", ``>=`` otherwise (sai/sai.py:211-214)."""
    v = np.asarray(values, dtype=np.float64)
    thr = outlier_threshold(v, q)
    if thr is None:
        return np.zeros(v.shape, dtype=bool)
    with np.errstate(invalid="ignore"):
        return (v > thr) if column.startswith("U") else (v >= thr)
