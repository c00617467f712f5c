"""`sai outlier` mirror (N1) and the multi-rank host logic, on CPU.

world_size-2 gloo groups stand in for the NCCL groups used on the GPUs."""

import os
import socket
import warnings

import numpy as np
import pytest

import sai_oracle as orc
from helpers import GOLDEN


def test_outlier_example_goldens(tmp_path):
    """Byte-identical to the outlier tables the reference ships for its example
    (examples/results/both/*.0.9.outliers.tsv, made from the scores.tsv next to them)."""
    from sai_b200.outlier import outlier

    d = os.path.join(GOLDEN, "outlier_example")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        outlier(os.path.join(d, "scores.tsv"), str(tmp_path / "o"), 0.9)
    expected = sorted(f for f in os.listdir(d) if f.startswith("expected."))
    assert len(expected) == 12
    for f in expected:
        col = f[len("expected.") : -len(".0.9.outliers.tsv")]
        assert (tmp_path / f"o.{col}.0.9.outliers.tsv").read_text() == open(os.path.join(d, f)).read(), col


def test_outlier_reference_unit_test(tmp_path):  # reference tests/test_sai.py:154-173
    import pandas as pd

    from sai_b200.outlier import outlier

    outlier(os.path.join(GOLDEN, "outlier_q.scores"), str(tmp_path / "outliers"), 0.25)
    df = pd.read_csv(tmp_path / "outliers.Q.0.25.outliers.tsv", sep="\t")
    assert df["Q"].iloc[0] == 0.7
    outlier(os.path.join(GOLDEN, "outlier_q.scores"), str(tmp_path / "outliers"), 0.75)
    df = pd.read_csv(tmp_path / "outliers.Q.0.75.outliers.tsv", sep="\t")
    assert df["Q"].iloc[0] == 1.0 and df["Q"].iloc[1] == 1.0


def test_outlier_degenerate_columns(tmp_path):
    from sai_b200.outlier import outlier

    p = tmp_path / "s.tsv"
    p.write_text("Chrom\tStart\tEnd\tRef\tTgt\tSrc\tOutgroup\tN(Variants)\tU\tQ\n"
                 "1\t1\t10\tA\tB\tC\tNA\t5\t2\tnan\n1\t11\t20\tA\tB\tC\tNA\t5\t2\tnan\n")
    with pytest.warns(UserWarning):
        outlier(str(p), str(tmp_path / "o"), 0.99)
    for col in ("U", "Q"):  # one unique value / no numeric value -> header only (sai.py:195-207)
        assert (tmp_path / f"o.{col}.0.99.outliers.tsv").read_text().count("\n") == 1


def test_thresholds_match_oracle():
    from sai_b200.outlier import outlier_mask, threshold_from_histogram, threshold_from_values

    rng = np.random.default_rng(0)
    for _ in range(1000):
        n = int(rng.integers(1, 60))
        q = float(rng.choice([0.25, 0.5, 0.9, 0.95, 0.99, rng.random()]))
        u = rng.integers(0, 7, size=n).astype(float)
        u[rng.random(n) < 0.1] = np.nan
        exp = orc.outlier_threshold(u, q)
        assert threshold_from_values(u, q) == exp
        assert threshold_from_histogram(np.bincount(u[~np.isnan(u)].astype(int), minlength=1), q) == exp
        assert np.array_equal(outlier_mask(u, exp, "U"), orc.outlier_mask(u, q, "U"))
        f = rng.random(n)
        f[rng.random(n) < 0.2] = np.nan
        assert threshold_from_values(f, q) == orc.outlier_threshold(f, q)
        assert np.array_equal(outlier_mask(f, threshold_from_values(f, q), "Q"), orc.outlier_mask(f, q, "Q"))


# ---------------------------------------------------------------- world_size 2 (gloo)
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from sai_b200.distributed import run_sharded, shard_ranges
    from sai_b200.outlier import distributed_threshold

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        u_all = rng.integers(0, 9, size=1001).astype(float)
        q_all = rng.random(1001)
        q_all[rng.random(1001) < 0.3] = np.nan
        cut = 300  # uneven shards
        mine = slice(0, cut) if rank == 0 else slice(cut, None)
        res = {}
        for q in (0.5, 0.95, 0.99):
            res[("U", q)] = distributed_threshold(u_all[mine], q, "U")
            res[("Q", q)] = distributed_threshold(q_all[mine], q, "Q")
        res["empty"] = distributed_threshold(np.array([np.nan]) if rank == 0 else np.array([]), 0.9, "Q")
        res["const"] = distributed_threshold(np.array([3.0, 3.0]), 0.9, "U")

        class FakePre:  # records the shard it was given (cf. reference tests/multiprocessing fakes)
            def run(self, chr_name, start, end):
                from sai_b200.windows import chunk_windows

                return [{"chr_name": chr_name, "start": s, "end": e, "rank": rank} for s, e in chunk_windows(start, end, 10000, 5000)]

        items = run_sharded(FakePre(), "21", 2309, 48989, 10000, 5000)
        res["items"] = items
        from sai_b200.distributed import run_genome_sharded

        spans = {"1": (120, 81000), "2": (7, 9000), "X": (40000, 66000)}
        res["genome"] = run_genome_sharded({c: FakePre() for c in spans}, spans, 10000, 5000)
        res["ranges"] = shard_ranges(2309, 48989, 10000, 5000, world)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_rank_threshold_and_sharding():
    import torch.multiprocessing as mp

    from sai_b200.windows import split_genome

    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = {r: dict(out[r]) for r in range(world)}
    rng = np.random.default_rng(5)
    u_all = rng.integers(0, 9, size=1001).astype(float)
    q_all = rng.random(1001)
    q_all[rng.random(1001) < 0.3] = np.nan
    for q in (0.5, 0.95, 0.99):
        for r in range(world):  # identical on every rank and equal to the single-process value
            assert res[r][("U", q)] == orc.outlier_threshold(u_all, q)
            assert res[r][("Q", q)] == orc.outlier_threshold(q_all, q)
    assert res[0]["empty"] is None and res[1]["empty"] is None and res[0]["const"] is None
    # sharding: the reference's golden chunk ranges, and rank 0 sees all windows in genome order
    assert res[0]["ranges"] == [(1, 30000), (25001, 55000)]  # tests/generators/test_chunk_generator.py:39
    wins = split_genome([2309, 48989], 10000, 5000)
    assert [(it["start"], it["end"]) for it in res[0]["items"]] == wins
    assert [it["rank"] for it in res[0]["items"]] == [0] * 5 + [1] * 5
    assert [(it["start"], it["end"]) for it in res[1]["items"]] == wins[5:]
    # whole genome: the flattened (chromosome, window) list cut in two; rank 0 sees every window once, in genome order
    spans = {"1": (120, 81000), "2": (7, 9000), "X": (40000, 66000)}
    flat = [(c, w) for c in spans for w in split_genome(list(spans[c]), 10000, 5000)]
    got = [(it["chr_name"], (it["start"], it["end"])) for it in res[0]["genome"]]
    assert got == flat
    half = (len(flat) + 1) // 2
    assert [it["rank"] for it in res[0]["genome"]] == [0] * half + [1] * (len(flat) - half)
    assert [(it["chr_name"], (it["start"], it["end"])) for it in res[1]["genome"]] == flat[half:]
