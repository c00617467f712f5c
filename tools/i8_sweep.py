#!/usr/bin/env python
"""int8 pipeline (sai_engine_score_host_i8) knob sweep in ONE process: the bench workload's shape, data built once.

    SAI_B200_LIB=tools/bin/libsai_b200_exp.so python tools/i8_sweep.py [--sites 6000000]

Each line: the knobs, ms per call (best of 3 after a warm-up call), int8 GB/s, wire bytes.  Knobs that only the
experiments build reads (tools/README.md): SAI_PACK_AHEAD, SAI_I8_SLICE_MB, SAI_I8_RING, SAI_I8_BLOCK_TILES.
"""
import argparse
import ctypes
import json
import mmap
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sai_b200.encode import MatrixGenotypes, make_layout  # noqa: E402
from sai_b200.scoring import HostEngine  # noqa: E402


def hugepage_array(shape):
    """int8 array in an anonymous mapping with MADV_HUGEPAGE (transparent huge pages when the kernel grants them)."""
    n = int(np.prod(shape))
    size = (n + (2 << 20) - 1) // (2 << 20) * (2 << 20)
    m = mmap.mmap(-1, size + (2 << 20))
    addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
    off = (-addr) % (2 << 20)
    libc = ctypes.CDLL(None, use_errno=True)
    rc = libc.madvise(ctypes.c_void_p(addr + off), ctypes.c_size_t(size), 14)  # MADV_HUGEPAGE
    a = np.frombuffer(m, dtype=np.int8, count=n, offset=off).reshape(shape)
    return a, m, rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=bench.WORKLOAD["n_sites"])
    ap.add_argument("--quick", action="store_true", help="only zt / dense / zt with the default knobs")
    a = ap.parse_args()
    wl = dict(bench.WORKLOAD)
    S, n_ind = a.sites, list(wl["n_ind"])
    rng = np.random.default_rng(1)
    nb = min(S, 32768)
    f = rng.beta(0.2, 2.0, size=nb)  # mostly rare variants, like the bench's synthetic spectrum
    block = np.concatenate([rng.binomial(2, f[:, None], size=(nb, n)).astype(np.int8) for n in n_ind], axis=1)
    reps = (S + nb - 1) // nb
    g = np.ascontiguousarray(np.tile(block, (reps, 1))[:S])
    pos = bench.make_positions(S, wl["mean_gap"], wl["seed"])
    ws, we = bench.make_windows(pos, wl["win_len"], wl["win_step"])
    job = bench.make_job_for(wl)
    lay = make_layout(n_ind, list(wl["ploidy"]), [2, 2, 2])
    cols = np.cumsum([0] + n_ind)

    def mg_of(arr):
        return MatrixGenotypes(lay, S, pos, [arr[:, cols[p] : cols[p + 1]] for p in range(3)])

    thp = "?"
    try:
        thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
    except OSError:
        pass
    print(json.dumps({"sites": S, "int8_gb": g.nbytes / 1e9, "cpus": os.cpu_count(), "thp": thp}), flush=True)
    eng = HostEngine(0)
    cpus = os.cpu_count() or 1
    base = {"SAI_PACK_AHEAD": "4", "SAI_I8_SLICE_MB": "32", "SAI_I8_RING": "4", "SAI_I8_BLOCK_TILES": "32"}
    cfgs = [dict(wire="zt"), dict(wire="dense"), dict(wire="zt", SAI_PACK_AHEAD="2"), dict(wire="zt", SAI_PACK_AHEAD="8"),
            dict(wire="zt", SAI_PACK_AHEAD="16"), dict(wire="zt", SAI_I8_SLICE_MB="16"), dict(wire="zt", SAI_I8_SLICE_MB="64"),
            dict(wire="zt", SAI_I8_BLOCK_TILES="8"), dict(wire="zt", SAI_I8_BLOCK_TILES="128"), dict(wire="zt", SAI_I8_RING="8"),
            dict(wire="zt", threads=cpus // 2), dict(wire="zt", threads=cpus // 4), dict(wire="zt", threads=1 if cpus < 3 else 3),
            dict(wire="zt", huge=True), dict(wire="dense", huge=True), dict(wire="zt")]
    if a.quick:
        cfgs = [dict(wire="zt"), dict(wire="dense"), dict(wire="zt"), dict(wire="zt")]
    huge = None
    ref = None
    for cfg in cfgs:
        env = dict(base)
        env.update({k: v for k, v in cfg.items() if k.startswith("SAI_")})
        os.environ.update(env)
        eng.set_i8_wire(cfg["wire"])
        eng.set_host_threads(cfg.get("threads", 0))
        arr = g
        note = {}
        if cfg.get("huge"):
            if huge is None:
                huge, _keep, rc = hugepage_array(g.shape)
                huge[:] = g
                note["madvise_rc"] = rc
                try:
                    note["AnonHugePages_kB"] = int([ln for ln in open("/proc/meminfo") if ln.startswith("AnonHugePages")][0].split()[1])
                except Exception:
                    pass
            arr = huge
        mg = mg_of(arr)
        r = eng.score_arrays(mg, ws, we, [job])
        if ref is None:
            ref = r
        same = bool(np.array_equal(r.u, ref.u) and np.array_equal(r.q, ref.q, equal_nan=True))
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            eng.score_arrays(mg, ws, we, [job])
            best = min(best, time.perf_counter() - t0)
        print(json.dumps({**cfg, **note, "ms": round(best * 1e3, 1), "int8_gbps": round(g.nbytes / best / 1e9, 1),
                          "wire_mb": round(eng.i8_wire_bytes() / 1e6, 1), "same_results": same}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
