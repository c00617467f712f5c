#!/usr/bin/env python
"""BASELINE.json configs 3-5 (+ all seven statistics) on the GPU: parity-test cases at full size, not bench lines.
Lives under tests/ because it checks against the CPU oracle; it is a script, not collected by pytest.

    python tests/run_configs.py --config 3            # two sources, haploid + diploid, missing, both anc modes
    python tests/run_configs.py --config 5            # threshold / window sweep on a 20k-sample cohort (cached counts)
    python tests/run_configs.py --config 6            # all seven statistics incl. DD, missing calls
    torchrun --nproc-per-node N tests/run_configs.py --config 4   # 22 autosomes, 80 M sites, sharded by window range
                                                                  # + genome-wide `sai outlier` thresholds

Every run spot-checks GPU results against the CPU oracle on decoded slices of
the device-generated data and prints one JSON summary line (rank 0).
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

import sai_oracle as orc  # noqa: E402
from sai_b200 import _cabi  # noqa: E402
from sai_b200.encode import PackedGenotypes, make_layout, unpack_population  # noqa: E402
from sai_b200.scoring import DeviceScorer, make_job, synth_fill  # noqa: E402
from sai_b200.windows import split_genome, split_windows_ranges  # noqa: E402

# hg19 autosome lengths (Mb, rounded) -- proportions for config 4
HG19_MB = [249, 243, 198, 191, 181, 171, 159, 146, 141, 136, 135, 134, 115, 107, 103, 90, 81, 78, 59, 63, 48, 51]


def positions(n_sites, mean_gap, seed):
    rng = np.random.default_rng(seed)
    return np.cumsum(rng.geometric(1.0 / mean_gap, size=n_sites).astype(np.int64)).astype(np.int32)


def device_matrix(lay, n_sites, roles, seed, missing):
    nbytes = int(_cabi.load().sai_packed_bytes(C.byref(lay), n_sites))
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    synth_fill(lay, d, n_sites, roles, seed, missing)
    return d


def decode_slice(lay, d_packed, pos, tile0, n_tiles):
    pps = lay.pairs_per_site
    sl = d_packed[tile0 * pps * 256 : (tile0 + n_tiles) * pps * 256].cpu().numpy()
    sub_pos = pos[tile0 * 32 : (tile0 + n_tiles) * 32]
    pg = PackedGenotypes(lay, len(sub_pos), sub_pos, sl)
    return sub_pos, [unpack_population(pg, p).astype(np.int64) for p in range(lay.n_pops)]


def check_windows(res, j, wins, sub_pos, mats, ploidy, src_idx, u_kw, q_kw, anc, max_checks=12):
    """Oracle comparison for the windows that lie inside the decoded slice."""
    n = 0
    for i, (s, e) in enumerate(wins):
        if s < sub_pos[0] or e > sub_pos[-1]:
            continue
        keep = (sub_pos >= s) & (sub_pos <= e)
        args = (mats[0][keep], mats[1][keep], [mats[k][keep] for k in src_idx], ploidy[0], ploidy[1], [ploidy[k] for k in src_idx])
        eu = orc.u_statistic(*args, pos=sub_pos[keep], anc_allele_available=anc, **u_kw)
        eq = orc.q_statistic(*args, pos=sub_pos[keep], anc_allele_available=anc, **q_kw)
        assert res.nsnps[j, i] == keep.sum(), (i, res.nsnps[j, i], keep.sum())
        assert res.u[j, i] == eu["value"], (i, res.u[j, i], eu["value"])
        assert np.array_equal(res.u_positions(j, i), eu["cdd_pos"]), i
        if np.isnan(eq["value"]):
            assert np.isnan(res.q[j, i]), i
        else:
            assert res.q[j, i] == float(eq["value"]), (i, res.q[j, i], float(eq["value"]))
        assert np.array_equal(res.q_positions(j, i), np.asarray(eq["cdd_pos"], dtype=np.int32)), i
        n += 1
        if n >= max_checks:
            break
    return n


def dev_windows(wins):
    ws = torch.tensor([w[0] for w in wins], dtype=torch.int64, device="cuda")
    we = torch.tensor([w[1] for w in wins], dtype=torch.int64, device="cuda")
    return ws, we


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# --------------------------------------------------------------------------
def config3(args):
    """Two source populations with joint y thresholds, haploid and diploid, with missing genotypes."""
    S = args.sites or 1_000_000
    out = {"config": 3, "n_sites": S, "cases": []}
    y_sets = [[("=", 1.0), ("=", 1.0)], [("=", 1.0), ("=", 0.0)], [("=", 0.0), ("=", 1.0)], [(">=", 0.5), ("<=", 0.5)]]
    for ploidy, n_src in ((2, (2, 2)), (1, (1, 1))):
        lay = make_layout([1500, 1000, n_src[0], n_src[1]], [ploidy] * 4, [2] * 4)
        d_packed = device_matrix(lay, S, [0, 1, 2, 2], 20261018 + 3 + ploidy, 0.002)
        pos = positions(S, 41.5, 3)
        d_pos = torch.from_numpy(pos).cuda()
        wins = split_genome([int(pos[0]), int(pos[-1])], 50_000, 10_000)
        d_ws, d_we = dev_windows(wins)
        sub_pos, mats = decode_slice(lay, d_packed, pos, 5000, 192)
        for anc in (True, False):
            jobs, kws = [], []
            for ys in y_sets:
                u_kw = dict(w=0.05, x=0.2, y_list=ys)
                q_kw = dict(w=0.05, quantile=0.95, y_list=ys)
                jobs.append(make_job(0, 1, [2, 3], anc, u=u_kw, q=q_kw))
                kws.append((u_kw, q_kw))
            sc = DeviceScorer(lay, S, len(wins), len(jobs), cap_u=1 << 20, cap_q=1 << 20)
            ms = timed(lambda: sc.step(d_packed, d_pos, d_ws, d_we, jobs))
            res = sc.results()
            checked = sum(check_windows(res, j, wins, sub_pos, mats, [ploidy] * 4, [2, 3], *kws[j], anc) for j in range(len(jobs)))
            out["cases"].append(dict(ploidy=ploidy, anc=anc, jobs=len(jobs), ms_per_pass=ms, windows=len(wins),
                                     u_sum=[int(x) for x in res.u.sum(1)], q_windows=[int(np.isfinite(r).sum()) for r in res.q],
                                     oracle_windows_checked=checked))
    print(json.dumps(out))


# --------------------------------------------------------------------------
def config5(args):
    """Threshold / window sweep on a 20 000-sample cohort: one genotype pass caches (num, called); the 20 (w, x, y)
    parameter sets are flagged from the cache 8 per launch -- once, the masks do not depend on the window shape --
    and every (win-len, step) grid is scored from the same masks, 8 sets per window launch."""
    S = args.sites or 1_000_000
    n_ind = [12_000, 7_996, 4]
    lay = make_layout(n_ind, [2, 2, 2], [2, 2, 2])
    d_packed = device_matrix(lay, S, [0, 1, 2], 20261018 + 5, 0.0)
    pos = positions(S, 41.5, 5)
    d_pos = torch.from_numpy(pos).cuda()
    sets = [(w, x, y) for w in (0.01, 0.05, 0.1, 0.2, 0.5) for x in (0.0, 0.01) for y in (0.5, 1.0)]
    kws = [(dict(w=w, x=x, y_list=[("=", y)]), dict(w=w, quantile=0.95, y_list=[("=", y)])) for w, x, y in sets]
    jobs = [make_job(0, 1, [2], True, u=u, q=q) for u, q in kws]
    grid_win = [(L, st) for L in (10_000, 50_000, 100_000) for st in (5_000, 10_000, 50_000) if st <= L]
    batches = [(b, min(b + 8, len(jobs))) for b in range(0, len(jobs), 8)]
    masters = [DeviceScorer(lay, S, 0, b1 - b0) for b0, b1 in batches]
    t_counts = timed(lambda: masters[0].site_counts(d_packed), reps=3)
    for m in masters[1:]:
        m.num, m.called = masters[0].num, masters[0].called
    alg = S * (sum(n_ind) * 2 / 8 + 4)

    def flag_all():
        for m, (b0, b1) in zip(masters, batches):
            m.flags_from_counts(jobs[b0:b1])

    t_flags = timed(flag_all, reps=10)  # all 20 sets
    t_flags1 = timed(lambda: masters[0].flags_from_counts(jobs[:1]), reps=10)  # one set per launch, for comparison
    t_flags8 = timed(lambda: masters[0].flags_from_counts(jobs[:8]), reps=10)
    flag_all()
    sub_pos, mats = decode_slice(lay, d_packed, pos, 9000, 128)
    t_win, checked, u_total, n_sets = 0.0, 0, 0, 0
    for (L, st) in grid_win:
        wins = split_genome([int(pos[0]), int(pos[-1])], L, st)
        d_ws, d_we = dev_windows(wins)
        scs = [m.sibling(len(wins), cap_u=1 << 22, cap_q=1 << 23) for m in masters]

        def win_all():
            for sc, (b0, b1) in zip(scs, batches):
                sc.window_stats(d_pos, d_ws, d_we, jobs[b0:b1])

        t_win += timed(win_all, reps=3)
        for sc, (b0, b1) in zip(scs, batches):
            res = sc.results()
            for j in range(0, b1 - b0, 3):
                checked += check_windows(res, j, wins, sub_pos, mats, [2, 2, 2], [2], *kws[b0 + j], True, max_checks=4)
            u_total += int(res.u.sum())
        n_sets += len(jobs)
    counts_bytes = S * 8 * 3  # (num, called) int32 pairs of ref, tgt, src
    print(json.dumps(dict(config=5, n_sites=S, n_samples=sum(n_ind), parameter_sets=n_sets, flag_sets=len(jobs),
                          genotype_pass_ms=t_counts, genotype_pass_gbps_algorithmic=alg / t_counts / 1e6,
                          flags_ms_all_20_sets=t_flags, flags_ms_per_set=t_flags / len(jobs),
                          flags_ms_one_set_launch=t_flags1, flags_ms_eight_set_launch=t_flags8,
                          flags_counts_gbps_eight_set_launch=counts_bytes / t_flags8 / 1e6,
                          window_stats_ms_all_160=t_win, window_stats_ms_per_set=t_win / n_sets,
                          sweep_ms_total=t_counts + t_flags + t_win,
                          oracle_windows_checked=checked, u_total=u_total)))


# --------------------------------------------------------------------------
def config4(args):
    """Whole genome: 22 autosomes (80 M sites by default) sharded by contiguous window ranges over the ranks, one
    launch pair per rank (sai_b200.genome), then the genome-wide outlier thresholds with one all_gather_into_tensor
    + the device select.  This is bench.py's `strong` record run on its own (bench.py prints it in every run)."""
    import types

    import torch.distributed as dist

    import bench

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = dict(bench.WORKLOAD)
    lay = make_layout(list(wl["n_ind"]), list(wl["ploidy"]), [2, 2, 2])
    rec = bench.strong_record(types.SimpleNamespace(genome_sites=args.sites or 80_000_000, steps=10), wl, lay, rank, world,
                              local, barrier)
    if rank == 0:
        print(json.dumps(dict(config=4, n_gpus=world, **rec)))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------
def config6(args):
    """All seven statistics (U, Q, Danc, Dplus, df, fd, DD) on one chromosome-scale matrix with missing calls:
    per-kernel times and oracle spot checks (the DD sums are exact integers)."""
    from sai_b200.encode import negative_table
    from sai_b200.scoring import dd_values, four_pop_values

    S = args.sites or 1_000_000
    n_ind = [1500, 1000, 4, 2]
    lay = make_layout(n_ind, [2, 2, 2, 2], [2, 2, 2, 2])
    d_packed = device_matrix(lay, S, [0, 1, 2, 0], 20261018 + 6, 0.002)
    pos = positions(S, 41.5, 6)
    d_pos = torch.from_numpy(pos).cuda()
    wins = split_genome([int(pos[0]), int(pos[-1])], 50_000, 10_000)
    d_ws, d_we = dev_windows(wins)
    # the device generator only knows one missing code: give every missing call a raw value (-1 = "0/.",
    # -2 = "./.") and build the negative-value table from the decoded matrix
    pg_all = PackedGenotypes(lay, S, pos, d_packed.cpu().numpy())
    mats = [unpack_population(pg_all, p) for p in range(lay.n_pops)]  # int8
    rng = np.random.default_rng(6)
    for m in mats:
        neg = m < 0
        m[neg] = np.where(rng.random(int(neg.sum())) < 0.5, -1, -2)
    off, n_site, n_ind_, n_val = negative_table(mats)
    d_neg = [torch.from_numpy(a).cuda() for a in (n_site, n_ind_, n_val)]
    u_kw = dict(w=0.01, x=0.5, y_list=[("=", 1.0)])
    q_kw = dict(w=0.01, quantile=0.95, y_list=[("=", 1.0)])
    job = make_job(0, 1, [2], True, u=u_kw, q=q_kw)
    sc = DeviceScorer(lay, S, len(wins), 1, cap_u=1 << 20, cap_q=1 << 21)
    t = {}
    t["site_flags+counts"] = timed(lambda: sc.site_flags(d_packed, [job], with_counts=True))
    t["window_stats"] = timed(lambda: sc.window_stats(d_pos, d_ws, d_we, [job]))
    t["window_patterns"] = timed(lambda: sc.pattern_sums(d_pos, d_ws, d_we, 0, 1, 3, [2]))
    t["site_hist"] = timed(lambda: sc.site_hist(d_packed, [0, 1]))
    hist, missing = sc.site_hist(d_packed, [0, 1])
    t["window_dd"] = timed(lambda: sc.dd_sums(d_packed, d_pos, d_ws, d_we, hist, 0, 1, [2], off, *d_neg))
    res = sc.results()
    sums = sc.pattern_sums(d_pos, d_ws, d_we, 0, 1, 3, [2]).cpu().numpy()
    ref_sum, tgt_sum, err = sc.dd_sums(d_packed, d_pos, d_ws, d_we, hist, 0, 1, [2], off, *d_neg)
    assert int(err.item()) == 0
    assert missing.cpu().tolist() == [int(off[1] - off[0]), int(off[2] - off[1])]
    four = four_pop_values(sums)
    dd = dd_values(ref_sum.cpu().numpy(), tgt_sum.cpu().numpy(), n_ind[0], n_ind[1], [n_ind[2]])
    checked = 0
    for i in np.linspace(0, len(wins) - 1, 10).astype(int):
        s, e = wins[i]
        keep = (pos >= s) & (pos <= e)
        if not keep.any():
            continue
        sub = [m[keep].astype(np.int64) for m in mats]
        assert res.nsnps[0, i] == keep.sum()
        assert res.u[0, i] == orc.u_statistic(sub[0], sub[1], [sub[2]], 2, 2, [2], pos=pos[keep], anc_allele_available=True, **u_kw)["value"]
        eq = orc.q_statistic(sub[0], sub[1], [sub[2]], 2, 2, [2], pos=pos[keep], anc_allele_available=True, **q_kw)["value"]
        assert (np.isnan(eq) and np.isnan(res.q[0, i])) or res.q[0, i] == float(eq)
        assert float(dd[0][i]).hex() == float(orc.dd_statistic(sub[0], sub[1], [sub[2]])[0]).hex(), i
        ef = orc.four_pop_statistics(sub[0], sub[1], [sub[2]], 2, 2, [2], out_gts=sub[3], out_ploidy=2)
        for name in ("Danc", "Dplus", "df", "fd"):
            a, b = four[name][0][i], ef[name][0]
            assert (np.isnan(a) and np.isnan(b)) or a == b, (name, i, a, b)  # numpy's pairwise order: bit-exact
        checked += 1
    print(json.dumps(dict(config="all-statistics", n_sites=S, n_samples=sum(n_ind), windows=len(wins),
                          missing_calls=int(off[-1]), ms={k: round(v, 4) for k, v in t.items()},
                          oracle_windows_checked=checked)))


# --------------------------------------------------------------------------
def config7(args):
    """The genotype pass on 3- and 4-plane populations (ploidy 4; B = 3 is what the encoder picks, B = 4 what
    high ploidy or the flipped-missing quirk needs): counts against the oracle on a slice, time and bandwidth."""
    S = args.sites or 1_000_000
    n_ind = [1500, 1000, 4]
    out = {"config": "wide-planes", "n_sites": S, "cases": []}
    for bits in (2, 3, 4):
        ploidy = 2 if bits == 2 else 4
        lay = make_layout(n_ind, [ploidy] * 3, [bits] * 3)
        d_packed = device_matrix(lay, S, [0, 1, 2], 20261018 + 7 + bits, 0.002)
        pos = positions(S, 41.5, 7)
        d_pos = torch.from_numpy(pos).cuda()
        wins = split_genome([int(pos[0]), int(pos[-1])], 50_000, 10_000)
        d_ws, d_we = dev_windows(wins)
        u_kw = dict(w=0.01, x=0.5, y_list=[("=", 1.0)])
        q_kw = dict(w=0.01, quantile=0.95, y_list=[("=", 1.0)])
        job = make_job(0, 1, [2], True, u=u_kw, q=q_kw)
        sc = DeviceScorer(lay, S, len(wins), 1, cap_u=1 << 20, cap_q=1 << 21)
        t_counts = timed(lambda: sc.site_counts(d_packed))
        t_fused = timed(lambda: sc.site_flags(d_packed, [job]))
        sc.step(d_packed, d_pos, d_ws, d_we, [job])
        res = sc.results()
        sub_pos, mats = decode_slice(lay, d_packed, pos, 7000, 160)
        num, called = sc.site_counts(d_packed)
        num, called = num.cpu().numpy(), called.cpu().numpy()
        for p in range(3):
            en, ec = orc.site_counts(mats[p])
            assert np.array_equal(num[p, 7000 * 32 : 7160 * 32], en) and np.array_equal(called[p, 7000 * 32 : 7160 * 32], ec)
        checked = check_windows(res, 0, wins, sub_pos, mats, [ploidy] * 3, [2], u_kw, q_kw, True, max_checks=8)
        alg = S * (sum(n_ind) * bits / 8 + 4)
        out["cases"].append(dict(bits=bits, ploidy=ploidy, packed_gb=d_packed.numel() / 1e9, site_counts_ms=t_counts,
                                 site_flags_ms=t_fused, site_flags_gbps_algorithmic=alg / t_fused / 1e6,
                                 site_flags_gbps_packed=d_packed.numel() / t_fused / 1e6, oracle_windows_checked=checked))
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4, 5, 6, 7])
    ap.add_argument("--sites", type=int, default=None)
    ap.add_argument("--dump", action="store_true")
    a = ap.parse_args()
    {3: config3, 4: config4, 5: config5, 6: config6, 7: config7}[a.config](a)
