#!/usr/bin/env python
"""torchrun worker of tests/test_gpu_configs.py::test_outlier_thresholds_nccl_two_gpus (and of
tools/gpu scripts): every rank scores its window-range shard of the small test genome on its own
GPU, then the genome-wide `sai outlier` thresholds come from ONE NCCL all_gather_into_tensor +
the device select.  Rank 0 writes a JSON verdict to $SAI_NCCL_TEST_OUT."""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sai_oracle as orc  # noqa: E402
from genome_helpers import rank_rows, score_rank, small_genome  # noqa: E402
from sai_b200.encode import make_layout  # noqa: E402
from sai_b200.genome import shard_genome  # noqa: E402
from sai_b200.outlier import device_thresholds, distributed_threshold  # noqa: E402
from sai_b200.scoring import make_job  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    chroms = small_genome()
    lay = make_layout([150, 100, 4], [2, 2, 2], [2, 2, 2])
    job = make_job(0, 1, [2], True, u=dict(w=0.05, x=0.3, y_list=[("=", 1.0)]), q=dict(w=0.05, quantile=0.95, y_list=[("=", 1.0)]))
    all_wins = [ch["wins"] for ch in chroms]
    mine = shard_genome(all_wins, world)[rank]
    batch, res = score_rank(chroms, mine, lay, job)
    sc = batch.scorer
    # device-resident columns: U (NaN for empty windows, like the score file) and Q
    u = sc.u[0].to(torch.float64)
    u = torch.where(sc.nsnps[0] == 0, torch.full_like(u, float("nan")), u)
    cols = torch.stack([u, sc.q[0]])
    out = {}
    for q in (0.9, 0.99):
        thr = device_thresholds(cols, q)
        # the host-array interface over the same NCCL group must agree
        thr_host = [distributed_threshold(cols[c].cpu().numpy(), q, name) for c, name in enumerate(("U", "Q"))]
        out[str(q)] = (thr, thr_host)
    # rank 0 also scores everything alone: rows and thresholds of the sharded run must be identical
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank_rows(res, batch, mine), out))
    verdict = None
    if rank == 0:
        whole_pieces = shard_genome(all_wins, 1)[0]
        b1, r1 = score_rank(chroms, whole_pieces, lay, job)
        whole = rank_rows(r1, b1, whole_pieces)
        merged = {}
        for rows, _ in gathered:
            merged.update(rows)
        keys = sorted(whole)
        u_all = np.array([np.nan if whole[k][0] == 0 else whole[k][1] for k in keys], dtype=np.float64)
        q_all = np.array([float.fromhex(whole[k][2]) for k in keys], dtype=np.float64)
        ok_thr = True
        for q in (0.9, 0.99):
            want = [orc.outlier_threshold(u_all, q), orc.outlier_threshold(q_all, q)]
            for _, o in gathered:
                ok_thr = ok_thr and o[str(q)][0] == want and o[str(q)][1] == want
        verdict = dict(world=world, rows_match_unsharded=merged == whole, thresholds_equal_across_ranks=ok_thr,
                       thresholds={q: gathered[0][1][q][0] for q in gathered[0][1]}, windows=len(whole))
        verdict["ok"] = bool(verdict["rows_match_unsharded"] and ok_thr)
        path = os.environ.get("SAI_NCCL_TEST_OUT")
        if path:
            with open(path, "w") as f:
                json.dump(verdict, f)
        print(json.dumps(verdict))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not verdict["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
