#!/usr/bin/env python
"""Per-rank pinned host -> device copy bandwidth with 1..N ranks copying at the same time
(plain cudaMemcpyAsync, one per copy; CUDA events; ranks released together by a barrier).
Separates a box limit (host memory / PCIe root sharing) from an engine limit when the end-to-end
number stops scaling with the GPU count.

    torchrun --nproc-per-node N tools/h2d_probe.py [--mb 1024] [--reps 8]

Rank 0 prints one JSON line: per-rank GB/s (copying alone is the N = 1 run) and the aggregate.
Also reports D2H and, with --pageable, the pageable-source copy rate."""

import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=8)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.mb << 20
    src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    src.fill_(rank + 1)
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    back = torch.empty(n, dtype=torch.uint8, pin_memory=True)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return n * a.reps / (e0.elapsed_time(e1) / 1e3) / 1e9

    h2d = timed(lambda: dst.copy_(src, non_blocking=True))
    d2h = timed(lambda: back.copy_(dst, non_blocking=True))
    # 32 MB slices, as the engines issue them
    sl = 32 << 20
    h2d_sliced = timed(lambda: [dst[o : o + sl].copy_(src[o : o + sl], non_blocking=True) for o in range(0, n, sl)])
    vals = torch.tensor([h2d, d2h, h2d_sliced], dtype=torch.float64, device="cuda")
    allv = [vals.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, vals)
    if rank == 0:
        rows = [v.cpu().tolist() for v in allv]
        print(json.dumps({
            "ranks": world, "mb_per_copy": a.mb, "reps": a.reps,
            "h2d_gbps_per_rank": [round(r[0], 2) for r in rows], "h2d_gbps_aggregate": round(sum(r[0] for r in rows), 2),
            "d2h_gbps_per_rank": [round(r[1], 2) for r in rows], "d2h_gbps_aggregate": round(sum(r[1] for r in rows), 2),
            "h2d_32mb_slices_gbps_per_rank": [round(r[2], 2) for r in rows],
            "cpus": os.cpu_count(),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
