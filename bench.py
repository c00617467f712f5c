#!/usr/bin/env python
"""Benchmark of the U + Q95 sliding-window scoring path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one chromosome-scale batch:
packed genotypes (resident in HBM) -> per-window N(Variants), U, Q95 and the
candidate position lists.  Default workload = BASELINE.json configs[1]:
synthetic 1000G-scale chr1, 6 M biallelic sites x 2504 diploid individuals
(ref 1500 / tgt 1000 / src 4), win-len 50 kb, step 10 kb, U(w=0.01, x=0.5,
y "=1") + Q(w=0.01, q=0.95, y "=1"), ancestral alleles available.

`value`  : windows/s, whole job (all ranks), inputs resident in HBM, CUDA events.
`e2e`    : same metric through the host-buffer C-ABI call, H2D/D2H inside the
           timed region, starting from the reference-side representation: int8
           allele-sum matrices in pageable host memory, packed on the fly by
           host threads (sai_engine_score_host_i8: pack | copy | K1 pipelined).
           `e2e.prepacked_zt` / `e2e.prepacked_dense`: the same call on tiles
           packed beforehand (zero-suppressed wire format / dense tiles).
`roofline`: the genotype pass (K1) against the measured HBM copy bandwidth.
`cpu_baseline`: the CPU oracle (numpy restatement of the reference) on a
           bounded sample of the same workload, all host cores.

Multi-GPU (torchrun): every rank scores its own chromosome-scale shard (weak
scaling, no data-path collective); time = max over ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(
    name="synthetic-1000G-chr1",
    n_sites=6_000_000,
    n_ind=(1500, 1000, 4),  # ref / tgt / src, diploid
    ploidy=(2, 2, 2),
    mean_gap=41.5,  # GRCh37 chr1 249 250 621 bp / 6e6 sites
    win_len=50_000,
    win_step=10_000,
    w=0.01,
    x=0.5,
    y=("=", 1.0),
    quantile=0.95,
    anc=True,
    seed=20261018 + 2,
)


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------
# clocks during the timed region (pynvml; falls back to nvidia-smi once)
# --------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.001):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.period = period_s
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        try:
            self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            try:
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            for bit, name in self.REASONS.items():
                if mask & bit and name != "gpu_idle":
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            self._stop.wait(self.period)

    def start(self):
        if self.nv is None:
            return
        self._stop.clear()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        if self.nv is None or self._thread is None:
            return
        self._once()
        self._stop.set()
        self._thread.join()
        self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {
            "sm_mhz": float(np.median(self.samples)),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# --------------------------------------------------------------------------
# workload construction
# --------------------------------------------------------------------------
def make_positions(n_sites: int, mean_gap: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    gaps = rng.geometric(1.0 / mean_gap, size=n_sites).astype(np.int64)
    pos = np.cumsum(gaps)
    assert pos[-1] < 2**31 - 1
    return pos.astype(np.int32)


def make_windows(pos: np.ndarray, win_len: int, win_step: int):
    from sai_b200.windows import split_genome

    wins = split_genome([int(pos[0]), int(pos[-1])], win_len, win_step)
    ws = np.array([w[0] for w in wins], dtype=np.int64)
    we = np.array([w[1] for w in wins], dtype=np.int64)
    return ws, we


def make_job_for(wl):
    from sai_b200.scoring import make_job

    return make_job(
        0, 1, [2], wl["anc"],
        u=dict(w=wl["w"], x=wl["x"], y_list=[wl["y"]]),
        q=dict(w=wl["w"], quantile=wl["quantile"], y_list=[wl["y"]]),
    )


def device_unpack_i8(lay, d_packed, n_sites: int, chunk_tiles: int = 2048) -> np.ndarray:
    """Synthetic-data preparation (untimed): decodes the device-generated packed tiles into the
    reference-side representation -- one row-major int8 matrix [sites, all individuals], populations
    as column blocks, missing = -1 -- in ordinary pageable host memory.  torch elementwise ops on
    the device (bit extraction), chunked; 2-plane populations only (the bench workload)."""
    import torch

    n_ind = [lay.pop[p].n_samples for p in range(lay.n_pops)]
    assert all(lay.pop[p].bits == 2 for p in range(lay.n_pops))
    out = np.empty((n_sites, sum(n_ind)), dtype=np.int8)
    pps = lay.pairs_per_site
    n_tiles = (n_sites + 31) // 32
    words = d_packed.view(torch.int32).view(n_tiles, pps, 32, 2)  # [tile, pair, site, plane]
    shifts = torch.arange(32, dtype=torch.int32, device=d_packed.device)
    col0 = np.cumsum([0] + n_ind)
    for t0 in range(0, n_tiles, chunk_tiles):
        t1 = min(n_tiles, t0 + chunk_tiles)
        rows = min(n_sites, t1 * 32) - t0 * 32
        block = torch.empty((rows, sum(n_ind)), dtype=torch.int8, device=d_packed.device)
        for p in range(lay.n_pops):
            L = lay.pop[p]
            w = words[t0:t1, L.pair_off : L.pair_off + L.n_pairs]  # [nt, groups, 32 sites, 2]
            a = (w[..., 0].unsqueeze(-1) >> shifts) & 1  # [nt, groups, site, individual]
            b = (w[..., 1].unsqueeze(-1) >> shifts) & 1
            code = (a + 2 * b).to(torch.int8)
            code = torch.where(code == 3, torch.full_like(code, -1), code)
            code = code.permute(0, 2, 1, 3).reshape((t1 - t0) * 32, L.n_groups * 32)
            block[:, col0[p] : col0[p + 1]] = code[:rows, : L.n_samples]
        out[t0 * 32 : t0 * 32 + rows] = block.cpu().numpy()
    return out


K1_SOURCES = ("site_kernels.cu", "site_cond.cuh", "popcount.cuh", "common.cuh")


def k1_source_sha() -> str:
    """Stamp of the sources that define k_site: profiles/k1_traffic.json (the ncu DRAM-byte
    capture, tools/k1_traffic.py) is only quoted when it was taken from these very sources."""
    import hashlib

    h = hashlib.sha256()
    for name in K1_SOURCES:
        with open(os.path.join(ROOT, "sai_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def algorithmic_bytes(wl, n_windows: int) -> int:
    """S x (sum_pops N_pop x 2 bit / 8 + 4) + W x 20."""
    per_site = sum(wl["n_ind"]) * 2 / 8 + 4
    return int(wl["n_sites"] * per_site + n_windows * 20)


# --------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on the host cores
# --------------------------------------------------------------------------
# kind "reference": the UNMODIFIED reference package (oracle/_ref, pip-installed from
# /root/reference by oracle/build_ref.py at build() time; it travels with the snapshot) --
# WindowGenerator._window_generator + FeaturePreprocessor.run driven by the reference's own
# sai.multiprocessing.mp_pool over ChunkGenerator._split_windows_ranges chunks (oracle/ref_driver.py).
# kind "port": oracle/sai_oracle.py over a fork pool -- only when oracle/_ref is absent.
_CPU = {}


def _cpu_chunk(args):
    import sai_oracle as orc

    start, end = args
    d = _CPU
    pos = d["pos"]
    lo, hi = np.searchsorted(pos, start, "left"), np.searchsorted(pos, end, "right")
    sub = lambda m: {k: orc.PopData(pos[lo:hi], v[lo:hi]) for k, v in m.items()}  # region read
    return orc.score_chunk("1", start, end, d["win_len"], d["win_step"], sub(d["ref"]), sub(d["tgt"]), sub(d["src"]),
                           d["pc"], d["sc"], d["anc"])


def cpu_reference_setup(wl, sample_sites: int, seed: int):
    """Decodes `sample_sites` of the synthetic workload (generated by the same
    device generator, so the CPU and GPU arms see the same data: the sample is
    the first `sample_sites` sites of the GPU arm's chromosome) into the int64
    matrices the reference holds.  Returns the windows fully inside the sample."""
    import ctypes as C

    import torch

    from sai_b200 import _cabi
    from sai_b200.encode import PackedGenotypes, make_layout, unpack_population
    from sai_b200.scoring import synth_fill

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_driver

    lay = make_layout(list(wl["n_ind"]), list(wl["ploidy"]), [2, 2, 2])
    nbytes = int(_cabi.load().sai_packed_bytes(C.byref(lay), sample_sites))
    d_packed = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    synth_fill(lay, d_packed, sample_sites, [0, 1, 2], seed, 0.0)
    pos = make_positions(wl["n_sites"], wl["mean_gap"], seed)[:sample_sites]
    pg = PackedGenotypes(lay, sample_sites, pos, d_packed.cpu().numpy())
    del d_packed
    mats = [unpack_population(pg, p).astype(np.int64) for p in range(3)]  # int64, as the reference holds them
    op, y = wl["y"]
    ystr = f"{op}{y}"
    ploidies = {"ref": {"REF": 2}, "tgt": {"TGT": 2}, "src": {"SRC": 2}}
    stats = {"U": {"ref": {"REF": wl["w"]}, "tgt": {"TGT": wl["x"]}, "src": {"SRC": ystr}},
             "Q": {"ref": {"REF": wl["w"]}, "tgt": {"TGT": wl["quantile"]}, "src": {"SRC": ystr}}}
    kind, why_port = "port", "oracle/_ref absent"
    if ref_driver.available():
        try:
            ref_driver.make_classes()  # imports the installed reference behind the three stubs
            kind, why_port = "reference", None
        except Exception as e:  # e.g. a dependency of the reference missing on this box: keep the bench line alive
            why_port = f"oracle/_ref not importable: {type(e).__name__}: {e}"
    _CPU.update(pos=pos, ref={"REF": mats[0]}, tgt={"TGT": mats[1]}, src={"SRC": mats[2]},
                win_len=wl["win_len"], win_step=wl["win_step"], anc=wl["anc"], ploidies=ploidies, stats=stats,
                kind=kind, why_port=why_port)
    if _CPU["kind"] == "reference":
        ref_driver.set_data("bench", pos, _CPU["ref"], _CPU["tgt"], _CPU["src"])
    else:
        from sai_b200.configs import PloidyConfig, StatConfig

        _CPU.update(pc=PloidyConfig(ploidies), sc=StatConfig(stats))
    from sai_b200.windows import split_genome

    wins = split_genome([int(pos[0]), int(pos[-1])], wl["win_len"], wl["win_step"])
    return [w for w in wins if w[1] <= int(pos[-1])]  # only windows fully inside the sample


def cpu_reference_run(wins, n_windows: int, nproc: int):
    """Scores the first `n_windows` windows of the sample, split into contiguous
    window ranges by ChunkGenerator._split_windows_ranges (8 chunks per worker,
    the commented-out `num_chunks=num_workers * 8` of sai/sai.py:91), mapped over
    `nproc` processes by the reference's mp_pool (fork Pool.map, mp_pool.py:70-71;
    pool start-up is part of the call, as it is for a reference user).
    Returns (items, seconds)."""
    use = wins[:n_windows]
    n_chunks = max(1, min(len(use), nproc * 8))
    d = _CPU
    t0 = time.perf_counter()
    if d["kind"] == "reference":
        import ref_driver

        items = ref_driver.score_windows("bench", "1", use, n_chunks, nproc, d["win_len"], d["win_step"],
                                         d["ploidies"], d["stats"], d["anc"])
    else:
        import multiprocessing as mp

        from sai_b200.windows import split_windows_ranges

        chunks = split_windows_ranges(use, n_chunks)
        if nproc <= 1:
            parts = list(map(_cpu_chunk, chunks))
        else:
            with mp.get_context("fork").Pool(nproc) as pool:
                parts = pool.map(_cpu_chunk, chunks)
        items = [it for part in parts for it in part]
    return items, time.perf_counter() - t0


def cpu_sample_text(n_windows, sample_sites, cores):
    if _CPU["kind"] == "reference":
        how = ("the unmodified reference (oracle/_ref: WindowGenerator + FeaturePreprocessor.run) under its own "
               f"sai.multiprocessing.mp_pool, {cores} processes")
    else:
        how = f"oracle/sai_oracle.py (numpy port; {_CPU.get('why_port')}) over a {cores}-process fork pool"
    return (f"first {n_windows} windows ({sample_sites} sites decoded to int64) of the same workload, {how}, "
            f"8 window-range chunks per worker")


def compare_items_with_gpu(items, res, j=0):
    """Window by window: the CPU arm's items against the GPU results of the same windows (the
    sample is a prefix of the GPU arm's chromosome).  N(Variants), U and both candidate lists must
    be identical, |dQ| <= 1e-12 with NaN <-> NaN (the contract); Q bit-identity is reported too."""
    isnan = lambda v: isinstance(v, (float, np.floating)) and np.isnan(v)
    n_bad, q_exact, max_dq = 0, True, 0.0
    for i, it in enumerate(items):
        ok = int(it["nsnps"]) == int(res.nsnps[j, i])
        if it["nsnps"] > 0:  # an empty window is NaN / NaN on both sides by construction of the item
            ok = ok and not isnan(it["U"]) and int(it["U"]) == int(res.u[j, i])
            ok = ok and np.array_equal(np.asarray(it["cdd_pos"]["U"], dtype=np.int64), res.u_positions(j, i).astype(np.int64))
            if isnan(it["Q"]):
                ok = ok and bool(np.isnan(res.q[j, i])) and int(res.q_cnt[j, i]) == 0
            else:
                dq = abs(float(it["Q"]) - float(res.q[j, i]))
                max_dq = max(max_dq, dq) if dq == dq else float("inf")
                q_exact = q_exact and float(it["Q"]) == float(res.q[j, i])
                ok = ok and dq <= 1e-12
                ok = ok and np.array_equal(np.asarray(it["cdd_pos"]["Q"], dtype=np.int64), res.q_positions(j, i).astype(np.int64))
        n_bad += 0 if ok else 1
    return {"windows_compared": len(items), "mismatches": n_bad, "q_bit_identical": bool(q_exact), "max_abs_dq": max_dq}


# --------------------------------------------------------------------------
# strong scaling: BASELINE config 4 (whole genome, sharded by window range + outlier threshold)
# --------------------------------------------------------------------------
HG19_MB = [249, 243, 198, 191, 181, 171, 159, 146, 141, 136, 135, 134, 115, 107, 103, 90, 81, 78, 59, 63, 48, 51]


def strong_record(args, wl, lay, rank, world, local, barrier):
    """The north_star partition, measured in the same launch as the headline: 22 autosomes with
    hg19-proportional lengths (80 M sites x 2504, 51 GB packed), the flattened (chromosome, window)
    list cut into `world` contiguous ranges like ChunkGenerator._split_windows_ranges, every rank
    loading the sites of its own ranges (+ halo) and scoring them with ONE genotype pass and ONE
    window launch (sai_b200.genome); then the genome-wide `sai outlier` threshold of the U and Q
    columns: one all_gather_into_tensor over NVLink + the device select (no host copy).
    TOTAL work is fixed as N grows: `ms` (max over ranks, CUDA events) should fall as 1/N.  With
    N > 1, rank 0 also scores the whole genome alone right after, so the speed-up is measured
    inside one run on one box (`n1_ms`)."""
    import torch
    import torch.distributed as dist

    from sai_b200.genome import GenomeBatch, piece_site_range, shard_genome, synth_fill_piece
    from sai_b200.outlier import device_thresholds
    from sai_b200.windows import split_genome

    total = int(args.genome_sites)
    chrom_sites = [max(64, int(total * mb / sum(HG19_MB)) // 32 * 32) for mb in HG19_MB]
    L, st = wl["win_len"], wl["win_step"]
    chrom_pos = [make_positions(n, mb * 1e6 / n, 100 + c) for c, (n, mb) in enumerate(zip(chrom_sites, HG19_MB))]
    chrom_wins = [split_genome([int(p[0]), int(p[-1])], L, st) for p in chrom_pos]
    total_windows = sum(len(w) for w in chrom_wins)
    job = make_job_for(wl)
    K = max(3, min(args.steps, 10))

    def build(pieces):
        ranges = [piece_site_range(chrom_pos[p.chrom], chrom_wins[p.chrom], p) for p in pieces]
        wins = [chrom_wins[p.chrom][p.win_lo : p.win_hi] for p in pieces]
        n_w = sum(len(w) for w in wins)
        batch = GenomeBatch(lay, [hi - lo for lo, hi in ranges], wins, 1, cap_u=max(8 * n_w, 1 << 16), cap_q=max(32 * n_w, 1 << 18))
        for k, (p, (lo, hi)) in enumerate(zip(pieces, ranges)):
            synth_fill_piece(batch, k, lo, chrom_sites[p.chrom], [0, 1, 2], 777 + p.chrom, 0.0)
            batch.pos_view(k).copy_(torch.from_numpy(chrom_pos[p.chrom][lo:hi]))
        torch.cuda.synchronize()
        return batch

    def time_passes(batch, sync_ranks):
        for _ in range(3):
            batch.score([job])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sync_ranks:
            barrier()
        a.record()
        for _ in range(K):
            batch.score([job])
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / K

    shards = shard_genome(chrom_wins, world)
    mine = shards[rank]
    batch = build(mine)
    ms_local = time_passes(batch, True)
    sc = batch.scorer
    res_totals = sc.totals.cpu().numpy()
    assert res_totals[0, 0] <= sc.cap_u and res_totals[0, 1] <= sc.cap_q, "candidate buffers too small"
    # the score columns stay on the device: U (NaN for an empty window, like the score file) and Q
    u = sc.u[0].to(torch.float64)
    cols = torch.stack([torch.where(sc.nsnps[0] == 0, torch.full_like(u, float("nan")), u), sc.q[0]])
    max_len = max(sum(p.win_hi - p.win_lo for p in s_) for s_ in shards)  # known from the sharding: no size exchange
    thr = device_thresholds(cols, 0.99, max_len=max_len)  # warm-up (NCCL sets the collective up lazily)
    reps, t_thr = 5, 0.0
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        thr = device_thresholds(cols, 0.99, max_len=max_len)
        t_thr += time.perf_counter() - t0
    t_thr /= reps
    stats = torch.tensor([ms_local, 1e3 * t_thr, float(sc.u[0].sum().item()), float(torch.isfinite(sc.q[0]).sum().item()),
                          float(sum(batch.n_sites))], dtype=torch.float64, device="cuda")
    per_rank = [stats.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, stats)
    per_rank = [t.cpu().tolist() for t in per_rank]
    ms = max(r[0] for r in per_rank)
    out = {
        "workload": f"BASELINE config 4: 22 autosomes, {sum(chrom_sites)} sites x {sum(wl['n_ind'])} diploid, "
                    f"win {L}/{st}, U + Q{int(wl['quantile'] * 100)}, sharded by contiguous window ranges (+ halo) over {world} GPU(s)",
        "scaling": "strong", "total_sites": int(sum(chrom_sites)), "total_windows": int(total_windows), "steps": K,
        "ms": ms, "windows_per_s": total_windows / (ms / 1e3), "per_rank_ms": [r[0] for r in per_rank],
        "per_rank_sites": [int(r[4]) for r in per_rank], "launches_per_rank_per_step": 2,
        "threshold_ms": max(r[1] for r in per_rank), "threshold_quantile": 0.99,
        "thresholds": {"U": thr[0], "Q": thr[1]},
        "threshold_how": "one all_gather_into_tensor of the device-resident U / Q columns + sai_column_quantiles (device radix select); wall clock incl. the 64-byte result read",
        "u_total": sum(r[2] for r in per_rank), "windows_with_q": int(sum(r[3] for r in per_rank)),
    }
    del batch, sc, cols
    torch.cuda.empty_cache()
    if world > 1:  # the same genome on ONE GPU of the same box, same run
        n1 = None
        if rank == 0:
            whole = build(shard_genome(chrom_wins, 1)[0])
            n1 = time_passes(whole, False)
            wu = whole.scorer.u[0].to(torch.float64)
            wc = torch.stack([torch.where(whole.scorer.nsnps[0] == 0, torch.full_like(wu, float("nan")), wu), whole.scorer.q[0]])
            # single-GPU thresholds through the same kernel (no process group involved)
            o = torch.empty((2, 4), dtype=torch.float64, device="cuda")
            import ctypes as C

            from sai_b200 import _cabi

            _cabi.check(_cabi.load().sai_column_quantiles(wc.data_ptr(), 2, 1, 0, wc.shape[1], wc.shape[1], 0.99, o.data_ptr(),
                                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            thr1 = o[:, 0].cpu().tolist()
            out["n1_ms"] = n1
            out["speedup_vs_one_gpu"] = n1 / ms
            out["u_total_one_gpu"] = float(whole.scorer.u[0].sum().item())
            out["thresholds_equal_one_gpu"] = bool(thr1[0] == thr[0] and thr1[1] == thr[1])
            del whole
            torch.cuda.empty_cache()
        barrier()
    return out if rank == 0 else None


def host_info():
    """CPU model / core count / vector paths of the box: the end-to-end number is bound by the host."""
    from sai_b200 import _cabi

    model, mhz = "?", None
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name") and model == "?":
                model = ln.split(":", 1)[1].strip()
            elif ln.startswith("cpu MHz") and mhz is None:
                mhz = float(ln.split(":", 1)[1])
    except OSError:
        pass
    lib = _cabi.load()
    return {"cpus": os.cpu_count(), "model": model, "mhz": mhz, "pack_isa": lib.sai_pack_isa().decode(), "zt_isa": lib.sai_zt_isa().decode()}


# --------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sites", type=int, default=WORKLOAD["n_sites"], help="override the number of sites (debug)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--cpu-sample-sites", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (debug)")
    ap.add_argument("--no-strong", action="store_true", help="skip the whole-genome strong-scaling record (debug)")
    ap.add_argument("--genome-sites", type=int, default=80_000_000, help="sites of the whole-genome strong-scaling record")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    wl = dict(WORKLOAD)
    wl["n_sites"] = args.sites
    rank, world, local = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        return reference_arm(args, wl, rank, world)

    import torch
    import torch.distributed as dist

    from sai_b200 import _cabi
    from sai_b200.encode import PackedGenotypes, make_layout
    from sai_b200.scoring import DeviceScorer, HostEngine, synth_fill
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- build the shard of this rank (untimed) ----
    seed = wl["seed"] + 1000 * rank
    lay = make_layout(list(wl["n_ind"]), list(wl["ploidy"]), [2, 2, 2])
    S = wl["n_sites"]
    packed_bytes = int(_cabi.load().sai_packed_bytes(C.byref(lay), S))
    d_packed = torch.empty(packed_bytes, dtype=torch.uint8, device="cuda")
    synth_fill(lay, d_packed, S, [0, 1, 2], seed, 0.0)
    pos = make_positions(S, wl["mean_gap"], seed)
    ws, we = make_windows(pos, wl["win_len"], wl["win_step"])
    W = int(ws.shape[0])
    d_pos, d_ws, d_we = torch.from_numpy(pos).cuda(), torch.from_numpy(ws).cuda(), torch.from_numpy(we).cuda()
    job = make_job_for(wl)
    sc = DeviceScorer(lay, S, W, 1, cap_u=max(4 * W, 1 << 16), cap_q=max(16 * W, 1 << 18))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        sc.step(d_packed, d_pos, d_ws, d_we, [job])
    torch.cuda.synchronize()
    res = sc.results()  # also checks the candidate capacity
    u_total, q_finite = int(res.u.sum()), int(np.isfinite(res.q).sum())

    # ---- timed region: K steps, inputs resident in HBM (3.8 GB per step >> 126 MB L2) ----
    K = args.steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1a = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    k1b = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    ev0.record()
    for i in range(K):
        k1a[i].record()
        sc.site_flags(d_packed, [job])
        k1b[i].record()
        sc.window_stats(d_pos, d_ws, d_we, [job])
    ev1.record()
    barrier()
    clocks.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(k1a, k1b)]))
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / K
    value = world * W / (ms_per_step / 1e3)

    # ---- e2e: HOST buffers -> C-ABI host engine -> HOST results, copies inside the timed region ----
    # Headline `e2e` starts where the reference's data starts: one int8 matrix of per-individual allele
    # sums per population in ordinary (pageable) host memory -- what reshape_genotypes leaves behind
    # (utils.py:405-410), narrowed to int8.  sai_engine_score_host_i8 packs it with host threads slice
    # by slice into pinned staging buffers while earlier slices are on the wire and in the genotype pass.
    # `e2e.prepacked_*`: the same call on tiles packed (and zt-encoded) beforehand, for callers that
    # keep the packed matrix around (parameter sweeps over host-resident data, a packed on-disk cache).
    Ke = args.e2e_steps if args.e2e_steps is not None else max(3, min(K, 10))
    e2e = None
    if Ke > 0:
        from sai_b200.encode import MatrixGenotypes, compress, pack_populations

        # this rank's share of the host cores (the 8-GPU boxes of this pool have 4 vCPUs per GPU); the packers
        # issue the copies and launches themselves and the calling thread sleeps in the C call, so no core is set aside
        host_threads = max(1, (os.cpu_count() or 1) // world)
        host_threads = _env_int("SAI_BENCH_HOST_THREADS", host_threads) or host_threads  # tuning knob (tools/)
        t0 = time.perf_counter()
        h_i8 = device_unpack_i8(lay, d_packed, S)  # [S, 2504] int8, pageable (synthetic-data preparation, untimed)
        t_unpack = time.perf_counter() - t0
        h_packed = torch.empty(packed_bytes, dtype=torch.uint8, pin_memory=True)
        h_packed.copy_(d_packed)
        torch.cuda.synchronize()
        del d_packed
        torch.cuda.empty_cache()
        cols = np.cumsum([0] + list(wl["n_ind"]))
        mats = [h_i8[:, cols[p] : cols[p + 1]] for p in range(3)]  # column blocks of the one matrix, as a VCF parse leaves them
        mg = MatrixGenotypes(lay, S, pos, mats)
        # the host packer alone (all of this rank's threads), into a pinned buffer: must reproduce the device-generated tiles
        barrier()
        t0 = time.perf_counter()
        pg_chk = pack_populations(mats, list(wl["ploidy"]), pos, bits=[2, 2, 2], n_threads=host_threads,
                                  out=torch.empty(packed_bytes, dtype=torch.uint8, pin_memory=True).numpy())
        t_pack = time.perf_counter() - t0
        t0 = time.perf_counter()
        pack_populations(mats, list(wl["ploidy"]), pos, bits=[2, 2, 2], n_threads=host_threads, out=pg_chk.packed)
        t_pack = min(t_pack, time.perf_counter() - t0)  # second pass: output pages already touched
        pack_matches = bool(np.array_equal(pg_chk.packed, h_packed.numpy()))
        del pg_chk
        pg = PackedGenotypes(lay, S, pos, h_packed.numpy())
        h_zt = torch.empty(int(_cabi.load().sai_zt_bound(C.byref(lay), S)), dtype=torch.uint8, pin_memory=True)
        t0 = time.perf_counter()
        zt = compress(pg, n_threads=host_threads, out=h_zt.numpy())
        t_encode = time.perf_counter() - t0
        h_off = torch.empty(zt.tile_off.shape[0], dtype=torch.int64, pin_memory=True)
        h_off.numpy()[:] = zt.tile_off.view(np.int64)
        zt.tile_off = h_off.numpy().view(np.uint64)
        eng = HostEngine(local)
        eng.set_host_threads(host_threads)

        def run_wire(data, steps):
            r = eng.score_arrays(data, ws, we, [job])  # warm-up (allocates device and staging buffers)
            same = bool(np.array_equal(r.u, res.u) and np.array_equal(r.q, res.q, equal_nan=True)
                        and np.array_equal(r.nsnps, res.nsnps)
                        and all(np.array_equal(r.u_positions(0, i), res.u_positions(0, i)) for i in range(0, W, 97)))
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                r = eng.score_arrays(data, ws, we, [job])
            t = time.perf_counter() - t0
            if world > 1:
                tt = torch.tensor([t], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = float(tt.item())
            return r, same, t / steps

        Ki = max(2, Ke)  # ~0.1 s per step on 16 cores
        r2, same_i8, t_i8 = run_wire(mg, Ki)  # default wire: zt records built by the packers
        i8_wire_bytes = eng.i8_wire_bytes()
        eng.set_i8_wire("dense")
        _, same_i8_dense, t_i8_dense = run_wire(mg, Ki)
        eng.set_i8_wire("auto")
        _, same_zt, t_zt = run_wire(zt, Ke)
        _, same_dense, t_dense = run_wire(pg, Ke)
        small = pos.nbytes + ws.nbytes + we.nbytes
        d2h = int(r2.nsnps.nbytes + r2.u.nbytes + r2.q.nbytes + r2.q_cnt.nbytes + r2.u_start.nbytes + r2.q_start.nbytes
                  + r2.totals.nbytes + 4 * int(r2.totals.sum()))
        eng.close()
        # int8 -> zt records in one pass, host only (sai_zt_pack_i8: the block encoder the pipeline's packers run):
        # must reproduce pack + encode byte for byte; its time is the zt pipeline's host stage running alone
        from sai_b200.encode import compress_matrices
        h_zt2 = torch.empty(h_zt.numel(), dtype=torch.uint8, pin_memory=True)  # (after the timed legs: leaves their memory state alone)
        t_pack_zt = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            zt2 = compress_matrices(mg, n_threads=host_threads, out=h_zt2.numpy())
            t_pack_zt = min(t_pack_zt, time.perf_counter() - t0)
        pack_zt_matches = bool(np.array_equal(zt2.stream, zt.stream) and np.array_equal(zt2.tile_off, zt.tile_off))
        del zt2, h_zt2
        tp = torch.tensor([t_pack, t_pack_zt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        t_pack, t_pack_zt = float(tp[0].item()), float(tp[1].item())
        e2e = {
            "value": world * W / t_i8, "unit": "windows/s",
            "h2d_bytes_per_step": int(i8_wire_bytes + small), "d2h_bytes_per_step": d2h, "steps": Ki, "ms_per_step": 1e3 * t_i8,
            "matches_device_path": same_i8,
            "input": f"int8 per-individual allele sums, {h_i8.nbytes / 1e9:.2f} GB of pageable host memory per rank "
                     f"(the reference holds the same matrix as int64); packed on the fly by {host_threads} host threads "
                     f"({_cabi.load().sai_pack_isa().decode()} row packer), each tile zt-encoded while still in the packer's L1 "
                     f"({_cabi.load().sai_zt_isa().decode()} record encoder), records streamed into pinned 32 MB ring slots, "
                     "pipelined with the copy, the device-side decode and K1",
            "host_threads": host_threads, "host": host_info(),
            "wire": "zt records built by the packers" if i8_wire_bytes < packed_bytes else "dense tiles (no vector record encoder on this CPU)", "wire_ratio": packed_bytes / max(1, i8_wire_bytes),
            "int8_gbps": h_i8.nbytes / t_i8 / 1e9,
            "pack_alone_ms": 1e3 * t_pack, "pack_alone_gbps_int8": h_i8.nbytes / t_pack / 1e9, "pack_reproduces_device_tiles": pack_matches,
            "pack_alone_note": "dense packer alone (int8 -> dense tiles in a pinned buffer, no GPU work)",
            "pack_zt_alone_ms": 1e3 * t_pack_zt, "pack_zt_reproduces_encode": pack_zt_matches,
            "pack_zt_alone_note": "int8 -> zt records alone (sai_zt_pack_i8 into a pinned buffer, incl. its final gather copy; no GPU work)",
            "wire_alone_ms": 1e3 * t_dense,
            "pipeline_vs_slowest_stage": t_i8 / max(min(t_pack, t_pack_zt), t_zt),
            "dense_wire": {
                "value": world * W / t_i8_dense, "unit": "windows/s", "ms_per_step": 1e3 * t_i8_dense,
                "h2d_bytes_per_step": int(packed_bytes + small), "matches_device_path": same_i8_dense,
                "pipeline_vs_slowest_stage": t_i8_dense / max(t_pack, t_dense),
                "note": "same call with sai_engine_set_i8_wire(e, 1): dense tiles through the ring (data without a hom-ref majority)",
            },
            "prepacked_zt": {
                "value": world * W / t_zt, "unit": "windows/s", "h2d_bytes_per_step": int(zt.stream.nbytes + zt.tile_off.nbytes + small),
                "d2h_bytes_per_step": d2h, "steps": Ke, "ms_per_step": 1e3 * t_zt, "matches_device_path": same_zt,
                "wire": "zt: zero-suppressed tiles (lossless), expanded to dense tiles on the device",
                "wire_ratio": packed_bytes / max(1, zt.stream.nbytes), "host_encode_s": t_encode,
                "note": "pinned host buffers packed and zt-encoded before the clock starts; ratio depends on the site-frequency spectrum",
            },
            "prepacked_dense": {
                "value": world * W / t_dense, "unit": "windows/s", "h2d_bytes_per_step": int(packed_bytes + small),
                "d2h_bytes_per_step": d2h, "steps": Ke, "ms_per_step": 1e3 * t_dense, "matches_device_path": same_dense,
            },
            "synthetic_unpack_s": t_unpack,
        }
        del h_i8, mats, mg

    # ---- roofline of the dominant kernel (K1) ----
    peak, peak_src = measured_peaks()
    alg = algorithmic_bytes(wl, W)
    achieved = alg / (k1_ms / 1e3) / 1e9
    traffic, traffic_note = None, "no ncu capture for this build (tools/gpu/gpu_traffic.sh writes profiles/k1_traffic.json)"
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("n_sites") != S:
            traffic_note = f"capture is for {tj.get('n_sites')} sites"
        elif tj.get("k1_source_sha") != k1_source_sha():
            traffic_note = "stale: the k_site sources changed since the ncu capture; re-run tools/gpu/gpu_traffic.sh"
        else:
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = tj.get("source")
    except Exception:
        pass

    # ---- strong scaling of the north-star partition: whole genome sharded by window range ----
    strong = None
    if not args.no_strong:
        strong = strong_record(args, wl, lay, rank, world, local, barrier)

    # ---- CPU baseline (rank 0, N=1 only): the reference itself on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample_sites = args.cpu_sample_sites or min(S, 480_000)
        wins = cpu_reference_setup(wl, sample_sites, seed)
        n_w = len(wins)  # the whole sample: ~10-20 s of CPU work on 16 cores
        items, secs = cpu_reference_run(wins, n_w, cores)
        cpu = {
            "value": len(items) / secs, "unit": "windows/s", "cores": cores, "kind": _CPU["kind"],
            "sample": cpu_sample_text(len(items), sample_sites, cores), "seconds": secs,
            # the same windows out of the GPU arm's device-resident pass, item by item
            "gpu_vs_cpu_arm": compare_items_with_gpu(items, res),
        }

    if rank == 0:
        out = {
            "metric": "u_q95_windows_per_sec",
            "value": value,
            "unit": "windows/s",
            "n_gpus": world,
            "steps": K,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u32 bit-planes (popcount) + f64 frequency compares",
            "data": "synthetic (device-generated, counter-based RNG)",
            "config": {
                "workload": f"{wl['name']}: {S} sites x {sum(wl['n_ind'])} diploid (ref {wl['n_ind'][0]}/tgt {wl['n_ind'][1]}/src {wl['n_ind'][2]}), "
                            f"win {wl['win_len']}/{wl['win_step']}, U(w={wl['w']},x={wl['x']},y={wl['y'][0]}{wl['y'][1]}) + Q{int(wl['quantile'] * 100)}",
                "windows_per_gpu": W,
                "sharding": "one chromosome-scale shard per GPU, no data-path collective",
                "l2": f"inputs ({packed_bytes / 1e9:.2f} GB per step) larger than L2 (126 MB); no explicit flush",
            },
            "genotype_gbps": world * alg / (ms_per_step / 1e3) / 1e9,
            "roofline": {
                "bound": "hbm", "kernel": "k_site (genotype pass, fused site conditions)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src, "k1_ms": k1_ms,
                "algorithmic_bytes": alg,
                "note": "peak is a copy (read + write) bandwidth; a read-only kernel with k_site's access pattern "
                        "reaches 7.0-7.2 TB/s on this GPU (tools/readbw.cu, profiles/round1_notes.md), so frac can exceed 1",
            },
            "cpu_baseline": cpu,
            "e2e": e2e,
            # the same record under the name VERDICT r1 asked for: e2e IS the from-int8 path now
            "e2e_from_int8": None if e2e is None else {k: e2e[k] for k in (
                "value", "unit", "ms_per_step", "pack_alone_ms", "wire_alone_ms", "pipeline_vs_slowest_stage", "host_threads")},
            "strong": strong,
            "gpu_launches": 2 * K,
            "clocks": clocks.summary(),
            "check": {"u_total": u_total, "windows_with_q": q_finite},
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def reference_arm(args, wl, rank, world):
    """The reference's CPU implementation of the path, timed on the host cores: the UNMODIFIED
    reference package (oracle/_ref, installed from /root/reference by oracle/build_ref.py; it
    travels to the GPU box with the snapshot) driven through its own mp_pool over all host cores
    on a bounded sample of the workload -- see cpu_reference_run.  Falls back to the numpy port
    (oracle/sai_oracle.py, kind "port") only when oracle/_ref is absent."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    S = wl["n_sites"]
    sample_sites = args.cpu_sample_sites or min(S, 240_000)
    wins = cpu_reference_setup(wl, sample_sites, wl["seed"])
    # calibrate: one short run, then size a step to fit the whole run in ~3 minutes
    n0 = min(len(wins), max(8, 2 * cores))
    items, secs = cpu_reference_run(wins, n0, cores)
    rate = len(items) / secs
    budget = 150.0 / (args.steps + args.warmup)
    n_w = int(max(n0, min(len(wins), rate * min(budget, 15.0))))
    for _ in range(args.warmup):
        cpu_reference_run(wins, n_w, cores)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        items, _s = cpu_reference_run(wins, n_w, cores)
        total += len(items)
    elapsed = time.perf_counter() - t0
    value = total / elapsed
    sample = "per step: " + cpu_sample_text(n_w, sample_sites, cores)
    out = {
        "impl": "reference",
        "metric": "u_q95_windows_per_sec",
        "value": value,
        "unit": "windows/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int64 allele sums + f64 (numpy)",
        "data": "synthetic (same device generator, decoded to int64 on the host)",
        "config": {"workload": f"{wl['name']}: bounded sample of {S} sites x {sum(wl['n_ind'])} diploid, "
                               f"win {wl['win_len']}/{wl['win_step']}, U + Q{int(wl['quantile'] * 100)}"},
        "cpu_baseline": {"value": value, "unit": "windows/s", "cores": cores, "kind": _CPU["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
