"""BASELINE.json configs 3, 4 and 5 under `pytest -m gpu`, against the oracle.

config 3  two source populations with joint y thresholds, haploid and diploid, missing genotypes,
          both ancestral-allele modes: the four y-sets as FUSED jobs of one genotype pass, 1 M sites.
config 4  whole genome: 22 chromosomes sharded 1/2/4/8-way by window range (halo, pieces side by
          side, one launch pair per rank) == unsharded == oracle, plus the genome-wide `sai outlier`
          thresholds from the device all-gather layout; the NCCL path itself with >= 2 GPUs.
config 5  threshold / window sweep on a 20 000-sample cohort through the cached counts
          (site_counts -> flags_from_counts, 8 parameter sets per launch -> window_stats): EVERY
          window of every (w, x, y) x (win-len, step) combination.

The oracle recomputes the per-site frequencies for every window (as the reference does); that is a
row-wise function, so these tests hand it a cache of `site_frequency` over the whole matrix
(`cached_site_frequency`; the identity slice-then-compute == compute-then-slice is asserted on the
fly).  Everything downstream of the frequencies -- validity, comparators, inversion, thresholds,
quantile, candidate lists -- runs in the oracle unchanged, per window.
"""

from __future__ import annotations

import contextlib
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import sai_oracle as orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- helpers
def _device_matrix(n_ind, ploidy, n_sites, roles, seed, missing):
    import torch

    from sai_b200 import _cabi
    from sai_b200.encode import make_layout
    from sai_b200.scoring import synth_fill

    lay = make_layout(list(n_ind), list(ploidy), [2] * len(n_ind))
    d = torch.empty(int(_cabi.load().sai_packed_bytes(C.byref(lay), n_sites)), dtype=torch.uint8, device="cuda")
    synth_fill(lay, d, n_sites, roles, seed, missing)
    return lay, d


def _decode(lay, d_packed, n_sites, pos):
    from sai_b200.encode import PackedGenotypes, unpack_population

    pg = PackedGenotypes(lay, n_sites, pos, d_packed.cpu().numpy())
    return [unpack_population(pg, p) for p in range(lay.n_pops)]  # int8, missing = -1


def _positions(n_sites, mean_gap, seed):
    rng = np.random.default_rng(seed)
    return np.cumsum(rng.geometric(1.0 / mean_gap, size=n_sites).astype(np.int64)).astype(np.int32)


@contextlib.contextmanager
def cached_site_frequency(mats, ploidy):
    """Lets the oracle look up `site_frequency(mats[k][a:b], ploidy[k])` in a per-matrix cache
    computed once by the oracle's own function (in row blocks).  Only exact row slices of the given
    matrices hit the cache; the first hits per matrix are re-computed the slow way and compared."""
    real = orc.site_frequency
    full, verified = [], [0] * len(mats)
    for m, p in zip(mats, ploidy):
        full.append(np.concatenate([real(m[i : i + 8192], p) for i in range(0, m.shape[0], 8192)]) if m.shape[0] else np.zeros(0))

    def lookup(gts, pl=1):
        for k, m in enumerate(mats):
            if isinstance(gts, np.ndarray) and gts.dtype == m.dtype and gts.ndim == 2 and gts.shape[1] == m.shape[1] \
                    and gts.strides == m.strides and pl == ploidy[k]:
                off = gts.ctypes.data - m.ctypes.data
                row = m.strides[0]
                if 0 <= off <= m.nbytes and off % row == 0 and off // row + gts.shape[0] <= m.shape[0] and \
                        (gts.shape[0] == 0 or np.shares_memory(gts, m)):
                    a = off // row
                    out = full[k][a : a + gts.shape[0]].copy()  # matching_loci writes into its frequencies
                    if verified[k] < 3 and gts.shape[0]:
                        verified[k] += 1
                        assert np.array_equal(out, real(gts, pl), equal_nan=True)
                    return out
        return real(gts, pl)

    orc.site_frequency = lookup
    try:
        yield
    finally:
        orc.site_frequency = real


def _check_window(res, j, i, lo, hi, pos, mats, ploidy, src_idx, u_kw, q_kw, anc):
    sub = [m[lo:hi] for m in mats]
    args = (sub[0], sub[1], [sub[k] for k in src_idx], ploidy[0], ploidy[1], [ploidy[k] for k in src_idx])
    assert res.nsnps[j, i] == hi - lo, (j, i)
    if hi == lo:  # empty window: the reference's item has NaN / NaN
        assert res.u[j, i] == 0 and np.isnan(res.q[j, i])
        return
    eu = orc.u_statistic(*args, pos=pos[lo:hi], anc_allele_available=anc, **u_kw)
    eq = orc.q_statistic(*args, pos=pos[lo:hi], anc_allele_available=anc, **q_kw)
    assert res.u[j, i] == eu["value"], (j, i, res.u[j, i], eu["value"])
    assert np.array_equal(res.u_positions(j, i), eu["cdd_pos"]), (j, i)
    if np.isnan(eq["value"]):
        assert np.isnan(res.q[j, i]), (j, i)
    else:
        assert res.q[j, i] == float(eq["value"]), (j, i, res.q[j, i], float(eq["value"]))  # bit-exact (<= 1e-12 is the contract)
    assert np.array_equal(res.q_positions(j, i), np.asarray(eq["cdd_pos"], dtype=np.int32)), (j, i)


def _window_bounds(pos, wins):
    ws = np.array([w[0] for w in wins])
    we = np.array([w[1] for w in wins])
    return np.searchsorted(pos, ws, "left"), np.searchsorted(pos, we, "right")


def _dev_windows(wins):
    import torch

    return (torch.tensor([w[0] for w in wins], dtype=torch.int64, device="cuda"),
            torch.tensor([w[1] for w in wins], dtype=torch.int64, device="cuda"))


# ---------------------------------------------------------------- config 5
def test_config5_sweep_every_window():
    """w in {0.01, 0.05, 0.1, 0.2, 0.5} x x in {0, 0.01} x y in {0.5, 1.0} (20 parameter sets) x
    win-len in {10, 50, 100 kb} x step in {5, 10, 50 kb} (step <= len: 8 shapes) on a
    20 000-sample cohort (ref 12 000 / tgt 7 996 / src 4), 32 768 sites: one genotype pass caches
    (num, called); the 20 sets are flagged from the cache 8 per launch and every window shape is
    scored from the same masks.  Every window's N, U, Q and both candidate lists == oracle."""
    import torch

    from sai_b200.scoring import DeviceScorer, make_job
    from sai_b200.windows import split_genome

    S = 32768
    n_ind, ploidy = [12_000, 7_996, 4], [2, 2, 2]
    lay, d_packed = _device_matrix(n_ind, ploidy, S, [0, 1, 2], 20261018 + 5, 0.0)
    pos = _positions(S, 41.5, 5)
    d_pos = torch.from_numpy(pos).cuda()
    sets = [(w, x, y) for w in (0.01, 0.05, 0.1, 0.2, 0.5) for x in (0.0, 0.01) for y in (0.5, 1.0)]
    kws = [(dict(w=w, x=x, y_list=[("=", y)]), dict(w=w, quantile=0.95, y_list=[("=", y)])) for w, x, y in sets]
    jobs = [make_job(0, 1, [2], True, u=u, q=q) for u, q in kws]
    shapes = [(L, st) for L in (10_000, 50_000, 100_000) for st in (5_000, 10_000, 50_000) if st <= L]
    assert len(sets) == 20 and len(shapes) == 8
    # one scorer holds the masks of all 20 sets; launches take 8 jobs at a time
    batches = [(b, min(b + 8, len(jobs))) for b in range(0, len(jobs), 8)]
    scorers = [DeviceScorer(lay, S, 0, b1 - b0) for b0, b1 in batches]
    num, called = scorers[0].site_counts(d_packed)
    for sc, (b0, b1) in zip(scorers, batches):
        sc.num, sc.called = num, called
        sc.flags_from_counts(jobs[b0:b1])
    # the cached-count path flags exactly what the fused genotype pass flags
    fused = DeviceScorer(lay, S, 0, 8)
    fused.site_flags(d_packed, jobs[:8])
    assert torch.equal(fused.mask_u, scorers[0].mask_u) and torch.equal(fused.mask_q, scorers[0].mask_q)
    mats = _decode(lay, d_packed, S, pos)
    n_windows = u_total = q_total = 0
    with cached_site_frequency(mats, ploidy):
        for L, st in shapes:
            wins = split_genome([int(pos[0]), int(pos[-1])], L, st)
            lo, hi = _window_bounds(pos, wins)
            d_ws, d_we = _dev_windows(wins)
            for master, (b0, b1) in zip(scorers, batches):
                sc = master.sibling(len(wins), cap_u=1 << 18, cap_q=1 << 19)  # same masks, this shape's result buffers
                sc.window_stats(d_pos, d_ws, d_we, jobs[b0:b1])
                res = sc.results()
                for j in range(b1 - b0):
                    u_kw, q_kw = kws[b0 + j]
                    for i in range(len(wins)):
                        _check_window(res, j, i, int(lo[i]), int(hi[i]), pos, mats, ploidy, [2], u_kw, q_kw, True)
                    n_windows += len(wins)
                    u_total += int(res.u[j, : len(wins)].sum())
                    q_total += int(np.isfinite(res.q[j, : len(wins)]).sum())
    assert n_windows == 20 * sum(len(split_genome([int(pos[0]), int(pos[-1])], L, st)) for L, st in shapes)
    assert u_total > 1000 and q_total > 1000


# ---------------------------------------------------------------- config 3
@pytest.mark.parametrize("ploidy_n, n_src", [(2, (2, 2)), (1, (1, 1))])
def test_config3_two_sources_fused_jobs_at_size(ploidy_n, n_src):
    """1 M sites x (1500 ref + 1000 tgt + two source populations), 0.2 % missing genotypes, the
    four joint y-sets {(=1,=1), (=1,=0), (=0,=1), (>=0.5,<=0.5)} as four fused jobs of ONE genotype
    pass, haploid and diploid, with and without ancestral alleles.  Every 4th window of every job
    (>= 1000 windows per job) against the oracle, plus tiling invariants over all windows."""
    import torch

    from sai_b200.scoring import DeviceScorer, make_job
    from sai_b200.windows import split_genome

    S = 1_000_000
    n_ind, ploidy = [1500, 1000, n_src[0], n_src[1]], [ploidy_n] * 4
    lay, d_packed = _device_matrix(n_ind, ploidy, S, [0, 1, 2, 2], 20261018 + 3 + ploidy_n, 0.002)
    pos = _positions(S, 41.5, 3)
    d_pos = torch.from_numpy(pos).cuda()
    wins = split_genome([int(pos[0]), int(pos[-1])], 50_000, 10_000)
    lo, hi = _window_bounds(pos, wins)
    d_ws, d_we = _dev_windows(wins)
    mats = _decode(lay, d_packed, S, pos)
    # the generator knows one missing code; the reference also sees -2 ("./."): same semantics (v < 0)
    rng = np.random.default_rng(33)
    for m in mats:
        neg = np.flatnonzero(m.reshape(-1) < 0)
        m.reshape(-1)[neg[rng.random(neg.size) < 0.5]] = -2
    y_sets = [[("=", 1.0), ("=", 1.0)], [("=", 1.0), ("=", 0.0)], [("=", 0.0), ("=", 1.0)], [(">=", 0.5), ("<=", 0.5)]]
    with cached_site_frequency(mats, ploidy):
        for anc in (True, False):
            kws = [(dict(w=0.05, x=0.2, y_list=ys), dict(w=0.05, quantile=0.95, y_list=ys)) for ys in y_sets]
            jobs = [make_job(0, 1, [2, 3], anc, u=u, q=q) for u, q in kws]
            sc = DeviceScorer(lay, S, len(wins), len(jobs), cap_u=1 << 21, cap_q=1 << 22)
            sc.step(d_packed, d_pos, d_ws, d_we, jobs)
            res = sc.results()
            flagged = np.unpackbits(sc.mask_u.cpu().numpy().view(np.uint8), bitorder="little").reshape(len(jobs), -1)
            for j in range(len(jobs)):
                for i in range(j % 4, len(wins), 4):
                    _check_window(res, j, i, int(lo[i]), int(hi[i]), pos, mats, ploidy, [2, 3], *kws[j], anc)
                # every site lies in exactly five 50 kb windows away from the chromosome ends
                inner = slice(5, len(wins) - 5)
                cs = np.concatenate([[0], np.cumsum(flagged[j, :S])])
                assert np.array_equal(res.u[j, inner], (cs[hi] - cs[lo])[inner])
            assert int(res.u.sum()) > 0 and int(np.isfinite(res.q).sum()) > 0


# ---------------------------------------------------------------- config 4
from genome_helpers import HG19_MB, rank_rows as _rows, score_rank as _score_rank, small_genome as _small_genome  # noqa: E402


def test_config4_sharded_genome_equals_unsharded_and_oracle():
    import torch

    from sai_b200 import _cabi
    from sai_b200.encode import make_layout
    from sai_b200.genome import shard_genome
    from sai_b200.outlier import device_thresholds
    from sai_b200.scoring import make_job

    chroms = _small_genome()
    lay = make_layout([150, 100, 4], [2, 2, 2], [2, 2, 2])
    u_kw = dict(w=0.05, x=0.3, y_list=[("=", 1.0)])
    q_kw = dict(w=0.05, quantile=0.95, y_list=[("=", 1.0)])
    job = make_job(0, 1, [2], True, u=u_kw, q=q_kw)
    all_wins = [ch["wins"] for ch in chroms]
    total = sum(len(w) for w in all_wins)

    # oracle: every window of every chromosome
    expect = {}
    for c, ch in enumerate(chroms):
        lo, hi = _window_bounds(ch["pos"], ch["wins"])
        with cached_site_frequency(ch["mats"], [2, 2, 2]):
            for i in range(len(ch["wins"])):
                sub = [m[lo[i] : hi[i]] for m in ch["mats"]]
                if hi[i] == lo[i]:
                    expect[(c, i)] = (0, 0, float("nan").hex(), [], [])
                    continue
                eu = orc.u_statistic(sub[0], sub[1], [sub[2]], 2, 2, [2], pos=ch["pos"][lo[i] : hi[i]], anc_allele_available=True, **u_kw)
                eq = orc.q_statistic(sub[0], sub[1], [sub[2]], 2, 2, [2], pos=ch["pos"][lo[i] : hi[i]], anc_allele_available=True, **q_kw)
                expect[(c, i)] = (int(hi[i] - lo[i]), int(eu["value"]), float(eq["value"]).hex(),
                                  [int(p) for p in eu["cdd_pos"]], [int(p) for p in eq["cdd_pos"]])
    assert len(expect) == total

    u_all = np.array([expect[k][1] for k in sorted(expect)], dtype=np.float64)
    q_all = np.array([float.fromhex(expect[k][2]) for k in sorted(expect)], dtype=np.float64)
    # an empty window is a NaN row in the score file (U too): feature_preprocessor.py:131-144
    empty = np.array([expect[k][0] == 0 for k in sorted(expect)])
    u_all[empty] = np.nan
    for world in (1, 2, 4, 8):
        shards = shard_genome(all_wins, world)
        assert sum(p.win_hi - p.win_lo for r in shards for p in r) == total
        got, per_rank = {}, []
        for pieces in shards:
            batch, res = _score_rank(chroms, pieces, lay, job)
            rows = _rows(res, batch, pieces)
            assert not (set(rows) & set(got))
            got.update(rows)
            u = res.u[0].astype(np.float64)
            u[res.nsnps[0] == 0] = np.nan
            per_rank.append(np.stack([u, res.q[0]]))
        assert got == expect, world  # rows identical to unsharded and to the oracle, bit for bit
        # genome-wide outlier thresholds from the all-gather layout [world, columns, max_len]
        max_len = max(a.shape[1] for a in per_rank)
        gathered = np.full((world, 2, max_len), np.nan)
        for r, a in enumerate(per_rank):
            gathered[r, :, : a.shape[1]] = a
        d = torch.from_numpy(gathered).cuda()
        out = torch.empty((2, 4), dtype=torch.float64, device="cuda")
        for q in (0.5, 0.9, 0.99, 1.0, 0.0):
            _cabi.check(_cabi.load().sai_column_quantiles(d.data_ptr(), 2, world, 2 * max_len, max_len, max_len, q,
                                                          out.data_ptr(), None))
            o = out.cpu().numpy()
            for col, vals in enumerate((u_all, q_all)):
                want = orc.outlier_threshold(vals, q)
                assert want is not None and o[col, 0] == want, (world, q, col, o[col], want)
                assert o[col, 1] == np.count_nonzero(~np.isnan(vals))
        if world == 1:  # single-GPU entry of the same function
            thr = device_thresholds(torch.from_numpy(per_rank[0]).cuda(), 0.99)
            assert thr == [orc.outlier_threshold(u_all, 0.99), orc.outlier_threshold(q_all, 0.99)]


@pytest.mark.parametrize("n", [0, 1, 2, 33, 1000, 70001])
def test_column_quantiles_vs_numpy(n):
    """sai_column_quantiles against numpy.quantile / the oracle's outlier_threshold: negative
    values, NaNs, ties, constant and empty columns, several chunks."""
    import torch

    from sai_b200 import _cabi

    rng = np.random.default_rng(n)
    cols = [
        rng.normal(size=n),  # negative and positive
        rng.integers(0, 6, size=n).astype(np.float64),  # U-like counts, many ties
        np.where(rng.random(n) < 0.3, np.nan, rng.random(n)),  # Q-like with NaN rows
        np.full(n, 0.25),  # constant: no threshold
        np.full(n, np.nan),  # empty
        np.concatenate([[-0.0], np.zeros(max(n - 1, 0))])[:n],  # -0.0 == 0.0: one distinct value
    ]
    n_chunks = 3
    ln = (n + n_chunks - 1) // n_chunks if n else 0
    buf = np.full((n_chunks, len(cols), max(ln, 1)), np.nan)
    for c, v in enumerate(cols):
        for k in range(n_chunks):
            part = v[k * ln : (k + 1) * ln]
            buf[k, c, : part.size] = part
    d = torch.from_numpy(buf).cuda()
    out = torch.empty((len(cols), 4), dtype=torch.float64, device="cuda")
    for q in (0.0, 0.3, 0.5, 0.95, 0.99, 1.0):
        _cabi.check(_cabi.load().sai_column_quantiles(d.data_ptr(), len(cols), n_chunks, buf.shape[1] * buf.shape[2],
                                                      buf.shape[2], buf.shape[2] if n else 0, q, out.data_ptr(), None))
        o = out.cpu().numpy()
        for c, v in enumerate(cols):
            want = orc.outlier_threshold(v, q)
            if want is None:
                assert np.isnan(o[c, 0]), (n, q, c, o[c])
            else:
                assert o[c, 0] == want, (n, q, c, o[c, 0], want)
            assert o[c, 1] == np.count_nonzero(~np.isnan(v))


def test_torch_ops_pieces_and_quantiles():
    """torch.ops.sai_b200.window_stats_pieces / column_quantiles (sai_b200/ops.py) == the ctypes path
    (GenomeBatch / sai_column_quantiles) on the small genome."""
    import torch

    import sai_b200.ops as ops
    from sai_b200.encode import make_layout
    from sai_b200.genome import shard_genome
    from sai_b200.scoring import make_job

    chroms = _small_genome(total_sites=20_000)
    lay = make_layout([150, 100, 4], [2, 2, 2], [2, 2, 2])
    job = make_job(0, 1, [2], True, u=dict(w=0.05, x=0.3, y_list=[("=", 1.0)]), q=dict(w=0.05, quantile=0.95, y_list=[("=", 1.0)]))
    pieces = shard_genome([ch["wins"] for ch in chroms], 1)[0]
    batch, res = _score_rank(chroms, pieces, lay, job)
    sc = batch.scorer
    z = lambda t: torch.zeros_like(t)
    nsnps, u, q, q_cnt, u_start, q_start, totals = z(sc.nsnps), z(sc.u), z(sc.q), z(sc.q_cnt), z(sc.u_start), z(sc.q_start), z(sc.totals)
    u_cand, q_cand = z(sc.u_cand), z(sc.q_cand)
    torch.ops.sai_b200.window_stats_pieces(batch.d_pos, batch.d_ws, batch.d_we, batch.d_first, batch.d_last, ops.jobs_tensor([job]),
                                           sc.mask_u, sc.mask_q, sc.qval, nsnps, u, q, q_cnt, u_start, q_start, totals, u_cand, q_cand)
    assert torch.equal(nsnps, sc.nsnps) and torch.equal(u, sc.u) and torch.equal(q_cnt, sc.q_cnt) and torch.equal(totals, sc.totals)
    assert np.array_equal(q.cpu().numpy(), res.q, equal_nan=True) and int(u.sum()) > 0
    cols = torch.stack([sc.u[0].to(torch.float64), sc.q[0]]).contiguous()
    out = torch.empty((2, 4), dtype=torch.float64, device="cuda")
    torch.ops.sai_b200.column_quantiles(cols, 0.9, out)
    o = out.cpu().numpy()
    assert o[0, 0] == orc.outlier_threshold(res.u[0].astype(np.float64), 0.9) and o[1, 0] == orc.outlier_threshold(res.q[0], 0.9)
    with pytest.raises(ValueError):
        torch.ops.sai_b200.column_quantiles(cols.to(torch.float32), 0.9, out)


def test_outlier_thresholds_nccl_two_gpus(tmp_path):
    """The NCCL path itself: two ranks score their shards of the small genome on their own GPUs
    and exchange ONE all_gather_into_tensor; thresholds identical on both ranks and equal to the
    oracle's over the unsharded rows."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "nccl.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", SAI_NCCL_TEST_OUT=str(out))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29621", os.path.join(ROOT, "tests", "nccl_genome_worker.py")]
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-4000:]
    r = json.loads(out.read_text())
    assert r["ok"] and r["world"] == 2 and r["thresholds_equal_across_ranks"] and r["rows_match_unsharded"], r
