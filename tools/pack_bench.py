#!/usr/bin/env python
"""Host packer throughput (int8 allele sums -> tiled bit-planes) by thread count and vector path.

    python tools/pack_bench.py [--sites 1000000]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sai_b200 import _cabi  # noqa: E402
from sai_b200.encode import make_layout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=1_000_000)
    ap.add_argument("--all-only", action="store_true", help="only the all-populations packer (the engine's path)")
    ap.add_argument("--threads", type=int, nargs="*", default=[1, 2, 4, 8, 16, 32, 64])
    a = ap.parse_args()
    lib = _cabi.load()
    S, n_ind = a.sites, [1500, 1000, 4]
    rng = np.random.default_rng(0)
    block = (rng.random((min(S, 65536), sum(n_ind)), dtype=np.float32) < 0.1).astype(np.int8)
    g = np.ascontiguousarray(np.tile(block, ((S + len(block) - 1) // len(block), 1))[:S])  # speed is data-independent
    lay = make_layout(n_ind, [2, 2, 2], [2, 2, 2])
    nbytes = int(lib.sai_packed_bytes(C.byref(lay), S))
    raw = np.zeros(nbytes + 4096, dtype=np.uint8)
    off = (-raw.ctypes.data) % 4096
    out = raw[off : off + nbytes]  # page aligned, like the engine's pinned staging ring
    cols = np.cumsum([0] + n_ind)
    res = {"isa_best": lib.sai_pack_isa().decode(), "nt_stores": os.environ.get("SAI_PACK_NT", "1") != "0", "knobs": {k: v for k, v in os.environ.items() if k.startswith("SAI_PACK_")}, "cpus": os.cpu_count(), "sites": S, "int8_gb": g.nbytes / 1e9, "gbps": {}}
    t0 = time.perf_counter(); g.copy(); res["numpy_copy_gbps_1thread"] = round(g.nbytes / (time.perf_counter() - t0) / 1e9, 2)
    for isa, name in (() if a.all_only else ((1, "portable"), (2, "sse2"), (0, "best"))):
        for th in (1, 2, 4, 8, 16, 32, 64):
            if th > (os.cpu_count() or 1):
                break
            best = 1e9
            for _ in range(2):
                t0 = time.perf_counter()
                for p in range(3):
                    rc = lib.sai_pack_i8_isa(C.byref(lay), p, g.ctypes.data + int(cols[p]), S, g.strides[0], out.ctypes.data, th, isa)
                    assert rc == 0
                best = min(best, time.perf_counter() - t0)
            res["gbps"][f"{name}_{th}"] = round(g.nbytes / best / 1e9, 2)
    ptrs = (C.c_void_p * 3)(*[g.ctypes.data + int(cols[p]) for p in range(3)])
    strides = (C.c_int64 * 3)(*[g.strides[0]] * 3)
    for th in a.threads:
        if th > (os.cpu_count() or 1):
            break
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            assert lib.sai_pack_i8_all(C.byref(lay), ptrs, strides, S, out.ctypes.data, th) == 0
            best = min(best, time.perf_counter() - t0)
        res["gbps"][f"all_pops_{th}"] = round(g.nbytes / best / 1e9, 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
