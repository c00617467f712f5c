"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle
and against the reference-generated golden fixtures.

Bars: U, N(Variants), candidate position lists bit-exact; Q bit-exact on every
fixture (the contract is |dQ| <= 1e-12 absolute, NaN <-> NaN -- asserted as
well so a regression shows which bar broke)."""

import json
import os

import numpy as np
import pytest

import sai_oracle as orc
import synth
from helpers import GOLDEN, SimplePloidy, SimpleStats, check_items, load_pipe_case, pipe_case_names, vcf_case_names

pytestmark = pytest.mark.gpu

Q_TOL = 1e-12  # absolute tolerance stated by BASELINE.json's north_star
FOUR_TOL = 0.0  # Danc/Dplus/df/fd: the pattern kernel sums in numpy's pairwise order -> bit-exact


@pytest.fixture(scope="module")
def engine():
    from sai_b200.scoring import HostEngine

    e = HostEngine(0)
    yield e
    e.close()


def _pop(pos, m):
    from sai_b200.encode import PopData

    return PopData(pos, m)


# ---------------------------------------------------------------- stat-class level
from test_oracle_golden import Q_KATS, U_KATS  # noqa: E402  (the reference's KATs)


@pytest.mark.parametrize("case", U_KATS)
def test_u_kat_gpu(case):
    from sai_b200.stats import STAT_REGISTRY

    ref, tgt, srcs, pl, w, x, y_list, anc, exp_u, exp_pos = case
    stat = STAT_REGISTRY.get("U")(ref_gts=np.array(ref), tgt_gts=np.array(tgt), src_gts_list=[np.array(s) for s in srcs],
                                  ref_ploidy=pl[0], tgt_ploidy=pl[1], src_ploidy_list=pl[2:])
    res = stat.compute(pos=np.arange(len(ref)), w=w, x=x, y_list=y_list, anc_allele_available=anc)
    assert res["name"] == "U" and res["value"] == exp_u
    assert np.array_equal(res["cdd_pos"], np.array(exp_pos))


@pytest.mark.parametrize("case", Q_KATS)
def test_q_kat_gpu(case):
    from sai_b200.stats import STAT_REGISTRY

    ref, tgt, srcs, pl, w, y_list, q, anc, exp_q, exp_pos = case
    stat = STAT_REGISTRY.get("Q")(ref_gts=np.array(ref), tgt_gts=np.array(tgt), src_gts_list=[np.array(s) for s in srcs],
                                  ref_ploidy=pl[0], tgt_ploidy=pl[1], src_ploidy_list=pl[2:])
    res = stat.compute(pos=np.arange(len(ref)), w=w, y_list=y_list, quantile=q, anc_allele_available=anc)
    assert res["name"] == "Q"
    if np.isnan(exp_q):
        assert np.isnan(res["value"])
    else:
        assert np.isclose(res["value"], exp_q)
    assert np.array_equal(res["cdd_pos"], np.array(exp_pos))


def test_q_edge_case_exact_gpu():
    from sai_b200.stats import QStatistic

    c = Q_KATS[4]
    res = QStatistic(ref_gts=np.array(c[0]), tgt_gts=np.array(c[1]), src_gts_list=[np.array(c[2][0])], ref_ploidy=1,
                     tgt_ploidy=1, src_ploidy_list=[1]).compute(
        pos=np.arange(3), w=0.95, y_list=[("=", 1.0)], quantile=0.95, anc_allele_available=False)
    assert float(res["value"]) == 0.9666666666666667


def test_stat_errors_gpu():
    from sai_b200.stats import QStatistic, UStatistic

    z = np.zeros((2, 2), int)
    kw = dict(ref_gts=z, tgt_gts=z, src_gts_list=[z], ref_ploidy=2, tgt_ploidy=2, src_ploidy_list=[2])
    with pytest.raises(ValueError, match="Missing required argument"):
        UStatistic(**kw).compute(pos=np.arange(2), w=0.5, x=0.5, y_list=[("=", 0)])
    with pytest.raises(ValueError, match="Missing required argument"):
        QStatistic(**kw).compute(pos=np.arange(2), w=0.5, quantile=0.95, anc_allele_available=False)
    with pytest.raises(ValueError, match=r"Parameters w must be within the range \[0, 1\]."):
        UStatistic(**kw).compute(pos=np.arange(2), w=1.1, x=0.5, y_list=[("=", 0)], anc_allele_available=True)
    with pytest.raises(ValueError, match="Invalid value in y_list"):
        UStatistic(**kw).compute(pos=np.arange(2), w=0.1, x=0.5, y_list=[("=", 1.5)], anc_allele_available=True)
    with pytest.raises(ValueError, match="Invalid operator in y_list"):
        UStatistic(**kw).compute(pos=np.arange(2), w=0.1, x=0.5, y_list=[("!", 0.5)], anc_allele_available=True)
    with pytest.raises(ValueError, match="The length of src_gts_list and y_list must match"):
        UStatistic(**kw).compute(pos=np.arange(2), w=0.1, x=0.5, y_list=[("=", 0.5)] * 2, anc_allele_available=True)
    with pytest.raises(ValueError, match="ploidy must be a positive integer"):
        UStatistic(**{**kw, "ref_ploidy": 0}).compute(pos=np.arange(2), w=0.1, x=0.5, y_list=[("=", 0.5)], anc_allele_available=True)


def test_stat_cases_gpu():
    """400 random cases whose expected outputs came from the reference's own
    UStatistic / QStatistic (tests/golden/make_golden.py)."""
    from sai_b200.stats import QStatistic, UStatistic

    meta = json.load(open(os.path.join(GOLDEN, "stat_cases.json")))
    arrs = np.load(os.path.join(GOLDEN, "stat_cases.npz"))
    for c, m in enumerate(meta):
        mats = [arrs[f"c{c}_g{k}"] for k in range(2 + m["n_src"])]
        pos = arrs[f"c{c}_pos"]
        y_list = [tuple(y) for y in m["y_list"]]
        pl = m["ploidy"]
        kw = dict(ref_gts=mats[0], tgt_gts=mats[1], src_gts_list=mats[2:], ref_ploidy=pl[0], tgt_ploidy=pl[1], src_ploidy_list=pl[2:])
        ru = UStatistic(**kw).compute(pos=pos, w=m["w"], x=m["x"], y_list=y_list, anc_allele_available=m["anc"])
        rq = QStatistic(**kw).compute(pos=pos, w=m["w"], quantile=m["q"], y_list=y_list, anc_allele_available=m["anc"])
        assert ru["value"] == m["U"], c
        assert [int(p) for p in ru["cdd_pos"]] == m["U_pos"], c
        if m["Q"] == "nan":
            assert np.isnan(rq["value"]), c
        else:
            assert abs(float(rq["value"]) - float.fromhex(m["Q"])) <= Q_TOL, c
            assert float(rq["value"]).hex() == m["Q"], c
        assert [int(p) for p in rq["cdd_pos"]] == m["Q_pos"], c


def test_dd_cases_gpu():
    """DD of the same 400 cases through the stat-class mirror: bit-exact against
    the reference's DdStatistic (integer sums on the GPU, the reference's own
    float64 steps on top)."""
    from sai_b200.stats import STAT_REGISTRY

    cls = STAT_REGISTRY.get("DD")
    meta = json.load(open(os.path.join(GOLDEN, "stat_cases.json")))
    arrs = np.load(os.path.join(GOLDEN, "stat_cases.npz"))
    for c, m in enumerate(meta):
        mats = [arrs[f"c{c}_g{k}"] for k in range(2 + m["n_src"])]
        pl = m["ploidy"]
        res = cls(ref_gts=mats[0], tgt_gts=mats[1], src_gts_list=mats[2:], ref_ploidy=pl[0], tgt_ploidy=pl[1],
                  src_ploidy_list=pl[2:]).compute()
        assert res["name"] == "DD" and [float(v).hex() for v in res["value"]] == m["DD"], c
    # tests/stats/test_dd_statistic.py:25-46
    r = cls(ref_gts=np.array([[1, 1], [0, 0]]), tgt_gts=np.array([[1, 0], [0, 1]]), src_gts_list=[np.array([[0, 1], [1, 1]])],
            ref_ploidy=1, tgt_ploidy=1, src_ploidy_list=[1]).compute()
    assert np.isclose(r["value"][0], 0.5)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_dd_random_vs_oracle(seed, engine):
    """DD sums against the oracle for shapes the fixtures do not reach: ploidy up to 8 (4 planes),
    missing values anywhere in -ploidy..-1, more than 8 and more than 32 source individuals
    (chunks / groups), several source populations, overlapping and empty windows."""
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import dd_values, make_job

    rng = np.random.default_rng(100 + seed)
    n_sites = int(rng.integers(50, 400))
    ploidy = [int(rng.choice([1, 2, 3, 4, 8])) for _ in range(4)]
    n_ind = [int(rng.integers(1, 90)), int(rng.integers(1, 70)), int(rng.choice([1, 3, 9, 33, 41])), int(rng.integers(1, 12))]
    f = rng.beta(0.4, 1.0, size=n_sites)
    mats = []
    for n, p in zip(n_ind, ploidy):
        g = rng.binomial(p, f[:, None], size=(n_sites, n)).astype(np.int8)
        miss = rng.random(g.shape) < float(rng.choice([0.0, 0.03, 0.3]))
        g[miss] = -rng.integers(1, p + 1, size=int(miss.sum()))
        mats.append(g)
    pos = np.cumsum(rng.integers(1, 30, size=n_sites)).astype(np.int32)
    top = int(pos[-1])
    wins = [(1, top), (int(pos[3]), int(pos[n_sites // 2])), (top + 1, top + 50), (int(pos[10]), int(pos[10]))]
    wins += [(s, s + 999) for s in range(1, top, 500)]
    pg = pack_populations(mats, ploidy, pos, keep_negatives=True)
    engine.score(pg, wins, [make_job(0, 1, [2, 3], True)])
    ref_sum, tgt_sum = engine.dd_sums(pg, 0, 1, [2, 3])
    got = dd_values(ref_sum, tgt_sum, n_ind[0], n_ind[1], n_ind[2:])
    m64 = [m.astype(np.int64) for m in mats]
    for i, (a, b) in enumerate(wins):
        keep = (pos >= a) & (pos <= b)
        if not keep.any():
            assert not ref_sum[:, i].any() and not tgt_sum[:, i].any()
            continue
        exp = orc.dd_statistic(m64[0][keep], m64[1][keep], [m64[2][keep], m64[3][keep]])
        for k in range(2):
            assert float(got[k][i]).hex() == float(exp[k]).hex(), (seed, i, k)
            assert np.array_equal(ref_sum[k, i, : n_ind[2 + k]],
                                  orc.cityblock_sums(m64[2 + k][keep], m64[0][keep]).sum(axis=1).astype(np.int64))


def test_torch_custom_ops_match_device_scorer(engine):
    """torch.ops.sai_b200.site_flags / window_stats (sai_b200/ops.py) give the results of the ctypes path."""
    import torch

    import sai_b200.ops as ops
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import make_job

    pos, mats = synth.make_populations(17, 5000, {"ref": {"R": (90, 2)}, "tgt": {"T": (70, 2)}, "src": {"S": (3, 2)}},
                                       mean_gap=60.0, introgressed=0.03, missing=0.01)
    g = [mats["ref"]["R"], mats["tgt"]["T"], mats["src"]["S"]]
    pg = pack_populations(g, [2, 2, 2], pos)
    wins = [(s, s + 19_999) for s in range(1, int(pos[-1]), 5_000)]
    jobs = [make_job(0, 1, [2], True, u=dict(w=0.05, x=0.2, y_list=[("=", 1.0)]), q=dict(w=0.05, quantile=0.95, y_list=[("=", 1.0)])),
            make_job(0, 1, [2], False, u=dict(w=0.3, x=0.1, y_list=[(">=", 0.5)]), q=dict(w=0.3, quantile=0.5, y_list=[(">=", 0.5)]))]
    ref = engine.score(pg, wins, jobs)
    dev = torch.device("cuda")
    J, W, nt = len(jobs), len(wins), pg.n_tiles
    packed = torch.from_numpy(pg.packed).to(dev)
    d_pos = torch.from_numpy(pg.pos).to(dev)
    ws = torch.tensor([w[0] for w in wins], dtype=torch.int64, device=dev)
    we = torch.tensor([w[1] for w in wins], dtype=torch.int64, device=dev)
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
    mask_u, mask_q, qval = z((J, nt), torch.int32), z((J, nt), torch.int32), z((J, nt * 32), torch.float64)
    nsnps, u, q, q_cnt = z((J, W), torch.int32), z((J, W), torch.int64), z((J, W), torch.float64), z((J, W), torch.int32)
    u_start, q_start, totals = z((J, W), torch.int64), z((J, W), torch.int64), z((J, 2), torch.int64)
    u_cand, q_cand = z((J, 4 * W + 1024), torch.int32), z((J, 4 * W + 1024), torch.int32)
    lay, jb = ops.struct_tensor(pg.layout), ops.jobs_tensor(jobs)
    torch.ops.sai_b200.site_flags(packed, lay, jb, pg.n_sites, mask_u, mask_q, qval, 0)
    torch.ops.sai_b200.window_stats(d_pos, ws, we, jb, mask_u, mask_q, qval, nsnps, u, q, q_cnt, u_start, q_start,
                                    totals, u_cand, q_cand)
    assert np.array_equal(nsnps.cpu().numpy(), ref.nsnps) and np.array_equal(u.cpu().numpy(), ref.u)
    assert np.array_equal(q.cpu().numpy(), ref.q, equal_nan=True) and np.array_equal(totals.cpu().numpy(), ref.totals)
    assert ref.u.sum() > 0
    with pytest.raises(ValueError):
        torch.ops.sai_b200.site_flags(packed, lay[:-1], jb, pg.n_sites, mask_u, mask_q, qval, 0)


def test_four_pop_stat_classes_gpu():
    """Danc / Dplus / df / fd through the registry (same constructor / compute() as the reference's
    classes), with and without an outgroup, against the oracle (1e-12 relative: summation order)."""
    from sai_b200.stats import STAT_REGISTRY

    assert sorted(STAT_REGISTRY.list_registered()) == sorted(["U", "Q", "DD", "Danc", "Dplus", "df", "fd"])
    rng = np.random.default_rng(42)
    for case in range(12):
        n = int(rng.integers(5, 300))
        pl = [int(rng.choice([1, 2, 4])) for _ in range(4)]
        f = rng.beta(0.5, 0.8, size=n)
        mk = lambda k, p: rng.binomial(p, f[:, None], size=(n, k)).astype(np.int64)
        ref, tgt, src1, src2, out = mk(30, pl[0]), mk(22, pl[1]), mk(2, pl[2]), mk(3, pl[2]), mk(2, pl[3])
        if case % 3 == 0:
            tgt[rng.random(tgt.shape) < 0.05] = -1
        with_out = case % 2 == 0
        kw = dict(ref_gts=ref, tgt_gts=tgt, src_gts_list=[src1, src2], ref_ploidy=pl[0], tgt_ploidy=pl[1],
                  src_ploidy_list=[pl[2], pl[2]], out_gts=out if with_out else None, out_ploidy=pl[3] if with_out else None)
        exp = orc.four_pop_statistics(ref, tgt, [src1, src2], pl[0], pl[1], [pl[2], pl[2]],
                                      out_gts=out if with_out else None, out_ploidy=pl[3] if with_out else None)
        for name in ("Danc", "Dplus", "df", "fd"):
            res = STAT_REGISTRY.get(name)(**kw).compute()
            assert res["name"] == name and len(res["value"]) == 2
            for a, b in zip(res["value"], exp[name]):
                assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= FOUR_TOL * max(1.0, abs(b)), (case, name, a, b)


def test_dd_needs_negative_table(engine):
    """The bit-planes keep one missing code; DD refuses to run without the raw
    values, and the engine checks that the table covers every missing call."""
    from sai_b200 import _cabi
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import make_job

    rng = np.random.default_rng(3)
    mats = [rng.integers(-2, 3, size=(200, n)).astype(np.int8) for n in (40, 33, 3)]
    pos = np.arange(1, 201, dtype=np.int32)
    pg = pack_populations(mats, [2, 2, 2], pos)
    engine.score(pg, [(1, 200)], [make_job(0, 1, [2], True)])
    with pytest.raises(ValueError, match="keep_negatives"):
        engine.dd_sums(pg, 0, 1, [2])
    pg = pack_populations(mats, [2, 2, 2], pos, keep_negatives=True)
    engine.score(pg, [(1, 200), (50, 120)], [make_job(0, 1, [2], True)])
    ref_sum, tgt_sum = engine.dd_sums(pg, 0, 1, [2])
    m64 = [m.astype(np.int64) for m in mats]
    for i, (a, b) in enumerate([(0, 200), (49, 120)]):
        assert np.array_equal(ref_sum[0, i], orc.cityblock_sums(m64[2][a:b], m64[0][a:b]).sum(axis=1).astype(np.int64))
        assert np.array_equal(tgt_sum[0, i], orc.cityblock_sums(m64[2][a:b], m64[1][a:b]).sum(axis=1).astype(np.int64))
    # a table that misses entries is rejected instead of silently dropping the calls
    lo, hi = int(pg.neg_off[0]), int(pg.neg_off[1])
    assert hi - lo > 3
    keep = np.ones(pg.neg_site.shape[0], bool)
    keep[lo + 1] = False
    pg.neg_site, pg.neg_ind, pg.neg_val = pg.neg_site[keep].copy(), pg.neg_ind[keep].copy(), pg.neg_val[keep].copy()
    pg.neg_off = pg.neg_off.copy()
    pg.neg_off[1:] -= 1
    with pytest.raises((_cabi.SaiError, ValueError), match="negative-value table"):
        engine.dd_sums(pg, 0, 1, [2])


# ---------------------------------------------------------------- pipeline level
@pytest.mark.parametrize("name", pipe_case_names())
def test_pipeline_golden_gpu(name, engine, tmp_path):
    from sai_b200.preprocessors import score_populations, write_items
    from sai_b200.windows import chunk_windows

    case, pos, data = load_pipe_case(name)
    wins = chunk_windows(case["start"], case["end"], case["win_len"], case["win_step"])
    mk = lambda d: {p: _pop(pos, m) for p, m in d.items()}
    stats = SimpleStats(case["stats"])
    items = score_populations(case["chr_name"], {t: wins for t in data["tgt"]}, mk(data["ref"]), mk(data["tgt"]),
                              mk(data["src"]), SimplePloidy(case["ploidies"]), stats, case["anc"], engine,
                              out_data=mk(data["outgroup"]) if "outgroup" in data else None)
    check_items(items, case["items"], q_tol=Q_TOL, four_tol=FOUR_TOL)
    check_items(items, case["items"], q_tol=0.0, four_tol=FOUR_TOL)  # U/Q bit-exact, so their TSV text is identical
    out = tmp_path / "scores.tsv"
    write_items(str(out), items, stats)
    assert out.read_text() == case["text"]["tsv"]  # every column, the four site-pattern statistics included
    for key in ("U", "Q"):
        if key in case["text"]:
            assert (tmp_path / f"scores.{key}.log").read_text() == case["text"][key]


@pytest.mark.parametrize("name", vcf_case_names())
def test_vcf_fixture_gpu(name, tmp_path):
    """VCF file -> ChunkPreprocessor.run -> process_items, against the outputs
    of the reference pipeline on the same records."""
    from sai_b200.preprocessors import ChunkPreprocessor

    case = json.load(open(os.path.join(GOLDEN, f"vcf_{name}.json")))
    stats = SimpleStats(case["stats"])
    out = tmp_path / "scores.tsv"
    pre = ChunkPreprocessor(
        vcf_file=os.path.join(GOLDEN, case["vcf"]),
        ref_ind_file=os.path.join(GOLDEN, f"vcf_{name}.ref.list"),
        tgt_ind_file=os.path.join(GOLDEN, f"vcf_{name}.tgt.list"),
        src_ind_file=os.path.join(GOLDEN, f"vcf_{name}.src.list"),
        out_ind_file=(os.path.join(GOLDEN, f"vcf_{name}.outgroup.list")
                      if os.path.exists(os.path.join(GOLDEN, f"vcf_{name}.outgroup.list")) else None),
        win_len=case["win_len"], win_step=case["win_step"], output_file=str(out),
        ploidy_config=SimplePloidy(case["ploidies"]), stat_config=stats,
        anc_allele_file=os.path.join(GOLDEN, f"vcf_{name}.anc.bed") if case["anc"] else None)
    items = pre.run(case["chr_name"], case["start"], case["end"])
    check_items(items, case["items"], q_tol=0.0, four_tol=FOUR_TOL)
    pre.process_items(items)
    if name == "outgroup_stats":  # reference tests/test_sai.py:92-110 (np.isclose against test.with.outgroup.res.tsv)
        want = dict(fd=0.0012826844929596443, df=0.0012417913767941238, Danc=-0.12498082112760132, Dplus=-0.12149240420484364)
        for k, v in want.items():
            assert np.isclose(items[0][k][0], v) and items[0][k][0] == v
    assert out.read_text() == case["text"]["tsv"]
    if name == "example_q":
        assert float(items[0]["Q"]) == 0.9  # reference tests/test_sai.py:63
    if name == "example_u":
        assert items[0]["U"] == 3  # reference tests/preprocessors/test_feature_preprocessor.py:223
    if name == "mixed_ploidy":
        assert [it["U"] for it in items] == [0, 1]  # reference tests/test_sai.py:150-151


def test_score_entry_point_gpu(tmp_path):
    """`score()` (signature of sai.sai.score): YAML config + VCF -> TSV + logs,
    Q == 0.9 like the reference's tests/test_sai.py:45-63."""
    import pandas as pd
    import yaml

    from sai_b200.score import score

    case = json.load(open(os.path.join(GOLDEN, "vcf_example_q.json")))
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(yaml.safe_dump({
        "statistics": case["stats"], "ploidies": case["ploidies"],
        "populations": {g: os.path.join(GOLDEN, f"vcf_example_q.{g}.list") for g in ("ref", "tgt", "src")}}))
    out = tmp_path / "sub" / "out.tsv"
    score(os.path.join(GOLDEN, case["vcf"]), "21", 6666, 6666, None, str(out), str(cfg), 1)
    df = pd.read_csv(out, sep="\t")
    assert list(df.columns) == ["Chrom", "Start", "End", "Ref", "Tgt", "Src", "Outgroup", "N(Variants)", "Q"]
    assert df["Q"].iloc[0] == 0.9
    assert out.read_text().split("\n", 1)[1] == case["text"]["tsv"]
    assert (tmp_path / "sub" / "out.Q.log").read_text() == "Chrom\tStart\tEnd\tQ_SNP\n" + case["text"]["Q"]
    with pytest.raises(FileNotFoundError, match="not found"):
        score(os.path.join(GOLDEN, case["vcf"]), "21", 6666, 6666, None, str(out), "config.yaml", 1)
    with pytest.raises(ValueError, match="not found in VCF"):
        score(os.path.join(GOLDEN, case["vcf"]), "7", 6666, 6666, None, str(out), str(cfg), 1)


def test_mp_pool_one_worker_per_gpu(tmp_path):
    """mp_pool (mirror of sai/multiprocessing/mp_pool.py): chunks of the ChunkGenerator scored by
    spawned worker processes, each with its own engine, give the file the serial loop writes."""
    import yaml

    from sai_b200.configs import load_config
    from sai_b200.generators import ChunkGenerator
    from sai_b200.multiprocessing import mp_pool
    from sai_b200.preprocessors import ChunkPreprocessor
    from sai_b200.score import score

    name = "outgroup_shape"
    case = json.load(open(os.path.join(GOLDEN, f"vcf_{name}.json")))
    lists = {g: os.path.join(GOLDEN, f"vcf_{name}.{g}.list") for g in ("ref", "tgt", "src")}
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(yaml.safe_dump({"statistics": case["stats"], "ploidies": case["ploidies"], "populations": lists}))
    anc = os.path.join(GOLDEN, f"vcf_{name}.anc.bed")
    vcf_path = os.path.join(GOLDEN, case["vcf"])
    serial = tmp_path / "serial.tsv"
    score(vcf_path, case["chr_name"], case["win_len"], case["win_step"], anc, str(serial), str(cfg), 3)
    c = load_config(str(cfg))
    pooled = tmp_path / "pooled.tsv"
    pre = ChunkPreprocessor(vcf_file=vcf_path, ref_ind_file=lists["ref"], tgt_ind_file=lists["tgt"], src_ind_file=lists["src"],
                            out_ind_file=None, win_len=case["win_len"], win_step=case["win_step"], output_file=str(pooled),
                            ploidy_config=c.ploidies, stat_config=c.statistics, anc_allele_file=anc)
    gen = ChunkGenerator(vcf_path, case["chr_name"], case["win_step"], case["win_len"], 3)
    assert len(gen) >= 2
    mp_pool(pre, gen, nprocess=2)
    rows = serial.read_text().split("\n", 1)[1]
    assert pooled.read_text() == rows and rows.count("\n") == len(case["items"])
    for key in ("U", "Q"):
        assert (tmp_path / f"pooled.{key}.log").read_text() == (tmp_path / f"serial.{key}.log").read_text().split("\n", 1)[1]


def test_score_mixed_ploidy_df_kat_gpu(tmp_path):
    """`score()` on the reference's mixed-ploidy VCF with its own config (tests/data/test_mixed_ploidy.config.yaml:
    df: True, fd: False, U with two sources) and its own assertions (tests/test_sai.py:127-151): no fd columns,
    df.src1 / df.src2 columns with the stored values, U rows [0, 1]."""
    import pandas as pd
    import yaml

    from sai_b200.score import score

    name = "mixed_ploidy"
    case = json.load(open(os.path.join(GOLDEN, f"vcf_{name}.json")))
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(yaml.safe_dump({
        "statistics": {"df": True, "fd": False, "U": case["stats"]["U"]}, "ploidies": case["ploidies"],
        "populations": {g: os.path.join(GOLDEN, f"vcf_{name}.{g}.list") for g in ("ref", "tgt", "src")}}, sort_keys=False))
    out = tmp_path / "scores.tsv"
    score(os.path.join(GOLDEN, case["vcf"]), "21", 50000, 50000, os.path.join(GOLDEN, f"vcf_{name}.anc.bed"), str(out), str(cfg), 1)
    df = pd.read_csv(out, sep="\t")
    assert "fd" not in df.columns and "fd.src1" not in df.columns and "fd.src2" not in df.columns
    assert list(df.columns) == ["Chrom", "Start", "End", "Ref", "Tgt", "Src", "Outgroup", "N(Variants)", "df.src1", "df.src2", "U"]
    assert np.isclose(df["df.src1"].iloc[0], -0.6086956521739131) and np.isclose(df["df.src2"].iloc[1], -0.45454545454545453)
    assert abs(df["df.src1"].iloc[0] - -0.6086956521739131) <= 1e-14 and abs(df["df.src2"].iloc[1] - -0.45454545454545453) <= 1e-14
    assert df["U"].iloc[0] == 0 and df["U"].iloc[1] == 1
    with pytest.raises(ValueError, match="requires polarized data"):  # tests/test_sai.py:113-124
        score(os.path.join(GOLDEN, case["vcf"]), "21", 50000, 50000, None, str(out), str(cfg), 1)


def test_source_combinations_gpu(engine):
    """num_src smaller than the number of source populations: one item list per source combination, outermost
    ref x tgt x combination like WindowGenerator (window_generator.py:164-167; the reference's
    tests/generators/test_window_generator.py:57-62 counts windows x 2 tgt x 2 combinations), each equal to
    scoring that combination on its own.  (The general path: the source ploidies are zipped positionally.)"""
    from sai_b200.preprocessors import score_populations
    from sai_b200.windows import chunk_windows

    pops = {"ref": {"R": (40, 2)}, "tgt": {"T1": (30, 2), "T2": (25, 2)}, "src": {"S1": (2, 2), "S2": (3, 2)}}
    pos, mats = synth.make_populations(23, 3000, pops, mean_gap=70.0, introgressed=0.03, missing=0.01)
    ploidies = {g: {p: v[1] for p, v in pops[g].items()} for g in pops}
    stats = SimpleStats({"U": {"ref": {"R": 0.1}, "tgt": {"T1": 0.2, "T2": 0.3}, "src": {"S1": "=1"}},
                         "Q": {"ref": {"R": 0.1}, "tgt": {"T1": 0.9, "T2": 0.5}, "src": {"S1": "=1"}}})
    wins = chunk_windows(1, int(pos[-1]) // 10000 * 10000 + 30000, 30000, 10000)
    mk = lambda d: {p: _pop(pos, m) for p, m in d.items()}
    both = score_populations("1", {t: wins for t in pops["tgt"]}, mk(mats["ref"]), mk(mats["tgt"]), mk(mats["src"]),
                             SimplePloidy(ploidies), stats, True, engine, num_src=1)
    assert len(both) == len(wins) * 2 * 2
    assert [(b["tgt_pop"], b["src_pop_list"]) for b in both[:: len(wins)]] == [("T1", ("S1",)), ("T1", ("S2",)), ("T2", ("S1",)), ("T2", ("S2",))]
    for t_i, t in enumerate(("T1", "T2")):
        for s_i, sname in enumerate(("S1", "S2")):
            alone = score_populations("1", {t: wins}, mk(mats["ref"]), {t: _pop(pos, mats["tgt"][t])},
                                      {sname: _pop(pos, mats["src"][sname])}, SimplePloidy(ploidies), stats, True, engine)
            part = both[(t_i * 2 + s_i) * len(wins) : (t_i * 2 + s_i + 1) * len(wins)]
            for a, b in zip(part, alone):
                assert (a["start"], a["nsnps"], a["U"]) == (b["start"], b["nsnps"], b["U"])
                assert (np.isnan(a["Q"]) and np.isnan(b["Q"])) or a["Q"] == b["Q"]
    assert sum(it["U"] for it in both if not np.isnan(it["U"])) > 0


def test_sharded_equals_unsharded_gpu(engine):
    """Window-range shards with their win_len - win_step halo (the multi-GPU
    partition, chunk_generator.py:111-142) give exactly the unsharded rows."""
    from sai_b200.preprocessors import score_populations
    from sai_b200.windows import chunk_windows, split_genome, split_windows_ranges

    pops = {"ref": {"R": (100, 2)}, "tgt": {"T": (80, 2)}, "src": {"S": (2, 2)}}
    stats = SimpleStats({"U": {"ref": {"R": 0.05}, "tgt": {"T": 0.2}, "src": {"S": "=1"}},
                         "Q": {"ref": {"R": 0.05}, "tgt": {"T": 0.9}, "src": {"S": "=1"}}})
    pc = SimplePloidy({"ref": {"R": 2}, "tgt": {"T": 2}, "src": {"S": 2}})
    pos, mats = synth.make_populations(41, 12000, pops, mean_gap=70.0, introgressed=0.02, missing=0.01)
    wins = split_genome([int(pos[0]), int(pos[-1])], 50000, 10000)

    def run(start, end):
        keep = (pos >= start) & (pos <= end)  # region read of the shard
        mk = lambda m: _pop(pos[keep], m[keep])
        return score_populations("3", {"T": chunk_windows(start, end, 50000, 10000)}, {"R": mk(mats["ref"]["R"])},
                                 {"T": mk(mats["tgt"]["T"])}, {"S": mk(mats["src"]["S"])}, pc, stats, True, engine)

    whole = run(*split_windows_ranges(wins, 1)[0])
    for n in (2, 3, 8):
        parts = [it for s, e in split_windows_ranges(wins, n) for it in run(s, e)]
        assert len(parts) == len(whole) == len(wins)
        for a, b in zip(parts, whole):
            assert (a["start"], a["end"], a["nsnps"], a["U"]) == (b["start"], b["end"], b["nsnps"], b["U"])
            assert (np.isnan(a["Q"]) and np.isnan(b["Q"])) or a["Q"] == b["Q"]
            assert np.array_equal(a["cdd_pos"]["U"], b["cdd_pos"]["U"]) and np.array_equal(a["cdd_pos"]["Q"], b["cdd_pos"]["Q"])
    assert sum(it["U"] for it in whole) > 0


# ---------------------------------------------------------------- K1 alone
@pytest.mark.parametrize("ploidy", [[2, 1, 4], [4, 8, 6], [3, 12, 2], [14, 5, 7], [30, 127, 60]])
@pytest.mark.parametrize("shape", [(1, [1, 1, 1]), (33, [5, 40, 2]), (1000, [257, 96, 3]), (4097, [1500, 1000, 4]),
                                   (300, [20000, 8, 1]), (2500, [256, 512, 129]), (700, [97, 1120, 1185])])
def test_site_counts_vs_oracle(ploidy, shape):
    """Counts of every population against calc_freq's numerator / called count for 2- to 8-plane
    populations (values up to 127: all of int8), with group counts that exercise the carry-save
    batches (8 pairs / 4 groups) and their remainders."""
    import torch

    from sai_b200.encode import pack_populations
    from sai_b200.scoring import DeviceScorer

    n_sites, n_ind = shape
    rng = np.random.default_rng(n_sites + 7 * ploidy[0])
    mats = []
    for n, p in zip(n_ind, ploidy):
        g = rng.integers(0, p + 1, size=(n_sites, n)).astype(np.int8)
        g[rng.random(g.shape) < 0.1] = -1
        g[rng.random(n_sites) < 0.05] = -2  # whole population missing at some sites
        mats.append(g)
    pg = pack_populations(mats, ploidy, np.arange(n_sites))
    want_bits = [next(b for b in range(2, 9) if (1 << b) - 1 > p) for p in ploidy]  # 2..8 planes
    assert [pg.layout.pop[i].bits for i in range(3)] == want_bits
    sc = DeviceScorer(pg.layout, n_sites, 0, 1)
    d_packed = torch.from_numpy(pg.packed).cuda()
    num, called = sc.site_counts(d_packed)
    num, called = num.cpu().numpy(), called.cpu().numpy()
    for i, g in enumerate(mats):
        en, ec = orc.site_counts(g)
        assert np.array_equal(num[i, :n_sites], en), i
        assert np.array_equal(called[i, :n_sites], ec), i
        assert not called[i, n_sites:].any()  # padding sites are all-missing


def test_site_variants_are_not_in_the_product_library():
    """The A/B variants of the genotype pass live in -DSAI_EXPERIMENTS builds only."""
    import torch

    from sai_b200.encode import pack_populations
    from sai_b200.scoring import DeviceScorer

    pg = pack_populations([np.zeros((40, 3), dtype=np.int8)], [2], np.arange(40))
    sc = DeviceScorer(pg.layout, 40, 0, 1)
    with pytest.raises(ValueError, match="SAI_EXPERIMENTS"):
        sc.site_counts(torch.from_numpy(pg.packed).cuda(), variant=5)


# ---------------------------------------------------------------- randomized differential
def _random_case(seed, n_sites, anc, missing, stats, pops, gap=60.0, win=(20000, 5000)):
    pos, mats = synth.make_populations(seed, n_sites, pops, mean_gap=gap, introgressed=0.02, missing=missing,
                                       src_all_missing=0.002 if missing else 0.0)
    ploidies = {g: {p: pops[g][p][1] for p in pops[g]} for g in pops}
    end = int(pos[-1]) // win[1] * win[1] + win[0]
    return pos, mats, ploidies, (1, end), win


@pytest.mark.parametrize("seed, missing, win", [(21, 0.0, (20000, 5000)), (22, 0.02, (3000, 3000)), (23, 0.3, (150000, 50000)),
                                                (24, 0.01, (700, 100))])
def test_all_statistics_differential_vs_oracle(seed, missing, win, engine):
    """All seven statistics, two targets x two sources + an outgroup (the fused multi-job path), random data with
    missing calls, windows from a handful of sites (< 8: numpy's sequential sum) to thousands (the pairwise recursion):
    every value equal to the oracle's, bit for bit."""
    from sai_b200.preprocessors import score_populations
    from sai_b200.windows import chunk_windows

    pops = {"ref": {"R": (120, 2)}, "tgt": {"T1": (90, 2), "T2": (40, 4)}, "src": {"N": (2, 2), "D": (3, 1)},
            "outgroup": {"O": (2, 2)}}
    stats = {"Danc": True, "DD": True, "fd": True,
             "U": {"ref": {"R": 0.05}, "tgt": {"T1": 0.3, "T2": 0.2}, "src": {"N": "=1", "D": ">=0.5"}},
             "df": True, "Dplus": True,
             "Q": {"ref": {"R": 0.1}, "tgt": {"T1": 0.95, "T2": 0.5}, "src": {"N": ">=0.5", "D": "=1"}}}
    pos, mats = synth.make_populations(seed, 12000, pops, mean_gap=25.0, introgressed=0.02, missing=missing,
                                       src_all_missing=0.002 if missing else 0.0)
    ploidies = {g: {p: pops[g][p][1] for p in pops[g]} for g in pops}
    start, end = 1, int(pos[-1]) // win[1] * win[1] + win[0]
    sc, pc = SimpleStats(stats), SimplePloidy(ploidies)
    wins = chunk_windows(start, end, *win)
    mkg = lambda d: {p: _pop(pos, m) for p, m in d.items()}
    got = score_populations("3", {t: wins for t in pops["tgt"]}, mkg(mats["ref"]), mkg(mats["tgt"]), mkg(mats["src"]), pc, sc,
                            True, engine, out_data=mkg(mats["outgroup"]))
    mk = lambda d: {p: orc.PopData(pos, m.astype(np.int64)) for p, m in d.items()}
    exp = orc.score_chunk("3", start, end, win[0], win[1], mk(mats["ref"]), mk(mats["tgt"]), mk(mats["src"]), pc, sc, True,
                          out_data=mk(mats["outgroup"]))
    assert len(got) == len(exp) == 2 * len(wins)
    same = lambda a, b: (np.isnan(a) and np.isnan(b)) or float(a).hex() == float(b).hex()
    n_val = 0
    for g, e in zip(got, exp):
        assert (g["start"], g["end"], g["tgt_pop"], g["nsnps"], g["out_pop"]) == (e["start"], e["end"], e["tgt_pop"], e["nsnps"], e["out_pop"])
        assert list(g.keys()) == list(e.keys())
        for s in ("U", "Q"):
            assert same(g[s], e[s]), (s, g["start"], g[s], e[s])
            assert np.array_equal(np.asarray(g["cdd_pos"][s]), np.asarray(e["cdd_pos"][s]))
        for s in ("Danc", "Dplus", "df", "fd", "DD"):
            gv, ev = (g[s], e[s]) if isinstance(e[s], list) else ([g[s]], [e[s]])
            assert len(gv) == len(ev) == 2
            for a, b in zip(gv, ev):
                assert same(a, b), (s, g["start"], a, b)
                n_val += not np.isnan(b)
    assert n_val > 20


@pytest.mark.parametrize("seed, anc, missing", [(11, True, 0.0), (12, False, 0.01), (13, False, 0.2), (14, True, 0.05)])
def test_differential_vs_oracle(seed, anc, missing, engine):
    from sai_b200.preprocessors import score_populations
    from sai_b200.windows import chunk_windows

    pops = {"ref": {"R": (300, 2)}, "tgt": {"T": (200, 2)}, "src": {"N": (2, 2), "D": (1, 2)}}
    stats = {"U": {"ref": {"R": 0.02}, "tgt": {"T": 0.3}, "src": {"N": "=1", "D": ">=0.5"}},
             "Q": {"ref": {"R": 0.1}, "tgt": {"T": 0.95}, "src": {"N": ">=0.5", "D": "=1"}}}
    pos, mats, ploidies, (start, end), win = _random_case(seed, 20000, anc, missing, stats, pops)
    sc, pc = SimpleStats(stats), SimplePloidy(ploidies)
    wins = chunk_windows(start, end, *win)
    got = score_populations("7", {"T": wins}, {p: _pop(pos, m) for p, m in mats["ref"].items()},
                            {p: _pop(pos, m) for p, m in mats["tgt"].items()},
                            {p: _pop(pos, m) for p, m in mats["src"].items()}, pc, sc, anc, engine)
    mk = lambda d: {p: orc.PopData(pos, m.astype(np.int64)) for p, m in d.items()}
    exp = orc.score_chunk("7", start, end, win[0], win[1], mk(mats["ref"]), mk(mats["tgt"]), mk(mats["src"]), pc, sc, anc)
    assert len(got) == len(exp) == len(wins)
    n_q = 0
    for g, e in zip(got, exp):
        assert (g["start"], g["end"], g["nsnps"]) == (e["start"], e["end"], e["nsnps"])
        for s in ("U", "Q"):
            if isinstance(e[s], float) and np.isnan(e[s]):
                assert np.isnan(g[s])
            elif s == "U":
                assert g[s] == e[s]
            else:
                n_q += 1
                assert abs(float(g[s]) - float(e[s])) <= Q_TOL
                assert float(g[s]) == float(e[s])
            assert np.array_equal(np.asarray(g["cdd_pos"][s]), np.asarray(e["cdd_pos"][s]))
    assert n_q > 10


def test_large_window_selection_paths(engine):
    """Windows with > 32 and > 1024 flagged sites exercise the shared-memory
    radix select and the unbuffered (global) select."""
    from sai_b200.preprocessors import score_populations

    pops = {"ref": {"R": (37, 2)}, "tgt": {"T": (211, 2)}, "src": {"S": (1, 2)}}
    stats = {"U": {"ref": {"R": 1.0}, "tgt": {"T": 0.0}, "src": {"S": ">=0"}},
             "Q": {"ref": {"R": 1.0}, "tgt": {"T": 0.95}, "src": {"S": ">=0"}}}
    pos, mats = synth.make_populations(21, 9000, pops, mean_gap=10.0, missing=0.15)
    wins = [(1, 400), (1, 3000), (1, 20000), (1, int(pos[-1])), (5000, 60000), (int(pos[-1]) + 5, int(pos[-1]) + 50)]
    ploidies = {g: {p: pops[g][p][1] for p in pops[g]} for g in pops}
    sc, pc = SimpleStats(stats), SimplePloidy(ploidies)
    for anc in (True, False):
        got = score_populations("1", {"T": wins}, {"R": _pop(pos, mats["ref"]["R"])}, {"T": _pop(pos, mats["tgt"]["T"])},
                                {"S": _pop(pos, mats["src"]["S"])}, pc, sc, anc, engine)
        for it, (s, e) in zip(got, wins):
            keep = (pos >= s) & (pos <= e)
            if not keep.any():
                assert it["nsnps"] == 0 and np.isnan(it["U"]) and np.isnan(it["Q"])
                continue
            sub = lambda m: m[keep].astype(np.int64)
            eu = orc.u_statistic(sub(mats["ref"]["R"]), sub(mats["tgt"]["T"]), [sub(mats["src"]["S"])], 2, 2, [2],
                                 pos=pos[keep], w=1.0, x=0.0, y_list=[(">=", 0.0)], anc_allele_available=anc)
            eq = orc.q_statistic(sub(mats["ref"]["R"]), sub(mats["tgt"]["T"]), [sub(mats["src"]["S"])], 2, 2, [2],
                                 pos=pos[keep], w=1.0, quantile=0.95, y_list=[(">=", 0.0)], anc_allele_available=anc)
            assert it["nsnps"] == int(keep.sum())
            assert it["U"] == eu["value"] and np.array_equal(it["cdd_pos"]["U"], eu["cdd_pos"])
            assert float(it["Q"]) == float(eq["value"])
            assert np.array_equal(it["cdd_pos"]["Q"], eq["cdd_pos"])
    assert max(it["nsnps"] for it in got) > 1024


# ---------------------------------------------------------------- device-resident path == host-buffer path
def test_device_scorer_matches_engine_and_cached_counts(engine):
    import torch

    from sai_b200.encode import pack_populations
    from sai_b200.scoring import DeviceScorer, make_job
    from sai_b200.windows import split_genome

    pops = {"ref": {"R": (150, 2)}, "tgt": {"T1": (90, 2), "T2": (70, 1)}, "src": {"S": (3, 2)}}
    pos, mats = synth.make_populations(31, 30000, pops, mean_gap=50.0, introgressed=0.02, missing=0.02)
    gts = [mats["ref"]["R"], mats["tgt"]["T1"], mats["tgt"]["T2"], mats["src"]["S"]]
    pg = pack_populations(gts, [2, 2, 1, 2], pos)
    wins = split_genome(pos, 50000, 10000)
    jobs = [
        make_job(0, 1, [3], True, u=dict(w=0.05, x=0.2, y_list=[("=", 1.0)]), q=dict(w=0.05, quantile=0.95, y_list=[("=", 1.0)])),
        make_job(0, 2, [3], False, u=dict(w=0.3, x=0.1, y_list=[(">=", 0.5)]), q=dict(w=0.5, quantile=0.5, y_list=[("<", 0.5)])),
    ]
    ref = engine.score(pg, wins, jobs)
    d_packed = torch.from_numpy(pg.packed).cuda()
    d_pos = torch.from_numpy(pg.pos).cuda()
    d_ws = torch.tensor([w[0] for w in wins], dtype=torch.int64, device="cuda")
    d_we = torch.tensor([w[1] for w in wins], dtype=torch.int64, device="cuda")
    for mode in ("fused", "counts"):
        sc = DeviceScorer(pg.layout, pg.n_sites, len(wins), len(jobs))
        if mode == "fused":
            sc.step(d_packed, d_pos, d_ws, d_we, jobs)
        else:
            sc.site_counts(d_packed)
            sc.flags_from_counts(jobs)
            sc.window_stats(d_pos, d_ws, d_we, jobs)
        got = sc.results()
        assert np.array_equal(got.nsnps, ref.nsnps) and np.array_equal(got.u, ref.u), mode
        assert np.array_equal(got.q, ref.q, equal_nan=True), mode
        assert np.array_equal(got.q_cnt, ref.q_cnt) and np.array_equal(got.totals, ref.totals), mode
        for j in range(len(jobs)):
            for i in range(len(wins)):
                assert np.array_equal(got.u_positions(j, i), ref.u_positions(j, i)), mode
                assert np.array_equal(got.q_positions(j, i), ref.q_positions(j, i)), mode
    assert ref.u.sum() > 0 and np.isfinite(ref.q).sum() > 10


def test_candidate_capacity_retry(engine):
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import make_job

    rng = np.random.default_rng(5)
    n = 5000
    gts = [np.zeros((n, 8), np.int8), rng.integers(1, 3, size=(n, 8)).astype(np.int8), np.full((n, 1), 2, np.int8)]
    pg = pack_populations(gts, [2, 2, 2], np.arange(1, n + 1))
    job = make_job(0, 1, [2], True, u=dict(w=0.5, x=0.1, y_list=[("=", 1.0)]), q=dict(w=0.5, quantile=0.0, y_list=[("=", 1.0)]))
    wins = [(1, n), (1, n // 2)]
    res = engine.score(pg, wins, [job], cap_u=10, cap_q=10)  # far too small: every site is a candidate
    assert res.u[0].tolist() == [n, n // 2]
    assert np.array_equal(res.u_positions(0, 0), np.arange(1, n + 1))
    assert np.array_equal(res.q_positions(0, 1), np.arange(1, n // 2 + 1))


# ---------------------------------------------------------------- synthetic generator + full-size properties
def _synth_setup(n_sites, n_ind=(1500, 1000, 4), missing=0.0, seed=20261019):
    import torch

    from sai_b200.encode import make_layout
    from sai_b200.scoring import synth_fill
    from sai_b200 import _cabi
    import ctypes as C

    lay = make_layout(list(n_ind), [2, 2, 2], [2, 2, 2])
    nbytes = int(_cabi.load().sai_packed_bytes(C.byref(lay), n_sites))
    d_packed = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    synth_fill(lay, d_packed, n_sites, [0, 1, 2], seed, missing)
    return lay, d_packed


def test_synthetic_subrange_vs_oracle():
    """Decodes slices of the device-generated matrix and checks the GPU results
    of the windows inside them against the oracle."""
    import torch

    from sai_b200.encode import PackedGenotypes, unpack_population
    from sai_b200.scoring import DeviceScorer, make_job

    n_sites = 64 * 1024
    for missing in (0.0, 0.01):
        lay, d_packed = _synth_setup(n_sites, (300, 200, 4), missing)
        pos = (np.arange(n_sites, dtype=np.int64) * 40 + 17).astype(np.int32)
        wins = [(s, s + 49999) for s in range(1, int(pos[-1]), 10000)]
        anc = missing == 0.0
        job = make_job(0, 1, [2], anc, u=dict(w=0.01, x=0.5, y_list=[("=", 1.0)]), q=dict(w=0.01, quantile=0.95, y_list=[("=", 1.0)]))
        sc = DeviceScorer(lay, n_sites, len(wins), 1)
        d_pos = torch.from_numpy(pos).cuda()
        d_ws = torch.tensor([w[0] for w in wins], dtype=torch.int64, device="cuda")
        d_we = torch.tensor([w[1] for w in wins], dtype=torch.int64, device="cuda")
        sc.step(d_packed, d_pos, d_ws, d_we, [job])
        got = sc.results()
        pg = PackedGenotypes(lay, n_sites, pos, d_packed.cpu().numpy())
        mats = [unpack_population(pg, p).astype(np.int64) for p in range(3)]
        assert got.u.sum() > 0
        for i in range(0, len(wins), 7):
            s, e = wins[i]
            keep = (pos >= s) & (pos <= e)
            eu = orc.u_statistic(mats[0][keep], mats[1][keep], [mats[2][keep]], 2, 2, [2], pos=pos[keep], w=0.01, x=0.5,
                                 y_list=[("=", 1.0)], anc_allele_available=anc)
            eq = orc.q_statistic(mats[0][keep], mats[1][keep], [mats[2][keep]], 2, 2, [2], pos=pos[keep], w=0.01,
                                 quantile=0.95, y_list=[("=", 1.0)], anc_allele_available=anc)
            assert got.nsnps[0, i] == keep.sum()
            assert got.u[0, i] == eu["value"]
            assert np.array_equal(got.u_positions(0, i), eu["cdd_pos"])
            assert (np.isnan(got.q[0, i]) and np.isnan(eq["value"])) or got.q[0, i] == float(eq["value"])
            assert np.array_equal(got.q_positions(0, i), np.asarray(eq["cdd_pos"], dtype=np.int32))


def test_full_size_properties():
    """BASELINE config 2 shape (6 M sites x 2504 diploid individuals): size-
    independent properties -- disjoint windows tile the genome (sum of U ==
    number of U-flagged sites == number of U candidates, sum of N(Variants) ==
    sites), overlapping 50 kb windows equal the sum of their five 10 kb parts,
    re-running is bit-identical, and a decoded slice matches the oracle."""
    import torch

    from sai_b200.encode import PackedGenotypes, unpack_population
    from sai_b200.scoring import DeviceScorer, make_job

    n_sites = 6_000_000
    lay, d_packed = _synth_setup(n_sites)
    rng = np.random.default_rng(3)
    pos = np.cumsum(rng.geometric(1 / 41.5, size=n_sites)).astype(np.int32)
    d_pos = torch.from_numpy(pos).cuda()
    job = make_job(0, 1, [2], True, u=dict(w=0.01, x=0.5, y_list=[("=", 1.0)]), q=dict(w=0.01, quantile=0.95, y_list=[("=", 1.0)]))
    last = int(pos[-1])
    small = [(s, s + 9999) for s in range(1, last + 1, 10000)]
    big = [(s, s + 49999) for s in range(1, last + 1, 10000)]

    def run(wins):
        sc = DeviceScorer(lay, n_sites, len(wins), 1, cap_u=400_000, cap_q=2_000_000)
        d_ws = torch.tensor([w[0] for w in wins], dtype=torch.int64, device="cuda")
        d_we = torch.tensor([w[1] for w in wins], dtype=torch.int64, device="cuda")
        sc.step(d_packed, d_pos, d_ws, d_we, [job])
        return sc, sc.results()

    sc_s, rs = run(small)
    flagged_u = int(np.unpackbits(sc_s.mask_u.cpu().numpy().view(np.uint8)).sum())
    assert int(rs.nsnps.sum()) == n_sites
    assert int(rs.u.sum()) == flagged_u == int(rs.totals[0, 0]) and flagged_u > 1000
    # disjoint windows: every U candidate appears exactly once, each window's list in genome order
    allc = np.sort(rs.u_cand[0, :flagged_u])
    assert np.all(np.diff(allc) > 0)
    for i in np.flatnonzero(rs.u[0] > 1)[:200]:
        assert np.all(np.diff(rs.u_positions(0, i)) > 0)
    _, rb = run(big)
    k = len(small)
    cs = np.concatenate([[0], np.cumsum(rs.u[0])])
    cn = np.concatenate([[0], np.cumsum(rs.nsnps[0].astype(np.int64))])
    hi = np.minimum(np.arange(k) + 5, k)
    assert np.array_equal(rb.u[0], cs[hi] - cs[np.arange(k)])
    assert np.array_equal(rb.nsnps[0], cn[hi] - cn[np.arange(k)])
    _, rb2 = run(big)
    assert np.array_equal(rb.q, rb2.q, equal_nan=True) and np.array_equal(rb.u, rb2.u)
    # decoded slices vs oracle: three places of the chromosome, every window inside them
    pps = lay.pairs_per_site
    checked = 0
    for t0, nt in ((1_000, 256), (100_000, 256), (180_000, 320)):
        sl = d_packed[t0 * pps * 256 : (t0 + nt) * pps * 256].cpu().numpy()
        sub_pos = pos[t0 * 32 : (t0 + nt) * 32]
        pg = PackedGenotypes(lay, nt * 32, sub_pos, sl)
        mats = [unpack_population(pg, p).astype(np.int64) for p in range(3)]
        here = 0
        for i, (s, e) in enumerate(big):
            if s < sub_pos[0] or e > sub_pos[-1]:
                continue
            keep = (sub_pos >= s) & (sub_pos <= e)
            eu = orc.u_statistic(mats[0][keep], mats[1][keep], [mats[2][keep]], 2, 2, [2], pos=sub_pos[keep], w=0.01, x=0.5,
                                 y_list=[("=", 1.0)], anc_allele_available=True)
            eq = orc.q_statistic(mats[0][keep], mats[1][keep], [mats[2][keep]], 2, 2, [2], pos=sub_pos[keep], w=0.01,
                                 quantile=0.95, y_list=[("=", 1.0)], anc_allele_available=True)
            assert rb.u[0, i] == eu["value"] and rb.nsnps[0, i] == keep.sum()
            assert np.array_equal(rb.u_positions(0, i), eu["cdd_pos"])
            assert (np.isnan(rb.q[0, i]) and np.isnan(eq["value"])) or rb.q[0, i] == float(eq["value"])
            assert np.array_equal(rb.q_positions(0, i), np.asarray(eq["cdd_pos"], dtype=np.int32))
            here += 1
        assert here >= 10
        checked += here
    assert checked >= 40


# ---------------------------------------------------------------- int8 pipeline (pack | copy | genotype pass)
def _same_results(a, b, n_jobs, n_windows):
    assert np.array_equal(a.nsnps, b.nsnps) and np.array_equal(a.u, b.u)
    assert np.array_equal(a.q, b.q, equal_nan=True) and np.array_equal(a.q_cnt, b.q_cnt)
    for j in range(n_jobs):
        for i in range(n_windows):
            assert np.array_equal(a.u_positions(j, i), b.u_positions(j, i)), (j, i)
            assert np.array_equal(a.q_positions(j, i), b.q_positions(j, i)), (j, i)


@pytest.mark.parametrize("n_sites, n_ind, ploidy, threads", [
    (1, [3, 2, 1], [2, 2, 2], 0), (33, [5, 40, 2], [2, 1, 2], 1), (5000, [257, 96, 3], [2, 1, 4], 3),
    (70_000, [1500, 1000, 4], [2, 2, 2], 0),  # 45 MB of tiles: three slices of the staging ring and a partial one
    (150_000, [1500, 1000, 4], [2, 2, 2], 5),  # more slices than ring slots: slots are reused
])
@pytest.mark.parametrize("wire", ["zt", "dense"])
def test_int8_pipeline_matches_packed_engine(n_sites, n_ind, ploidy, threads, wire, engine):
    """sai_engine_score_host_i8 (int8 matrices packed by host threads into the pinned ring while
    earlier slices are copied and flagged) == packing first and sai_engine_score_host, bit for bit;
    a follow-up batch of jobs over the resident tiles (sai_engine_score_resident) likewise.  Both
    wire formats: zt records built by the packers (default; decoded on the GPU) and dense tiles."""
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import make_job
    from sai_b200.windows import split_genome

    rng = np.random.default_rng(n_sites)
    f = rng.beta(0.3, 1.5, size=n_sites)
    whole = np.empty((n_sites, sum(n_ind)), dtype=np.int8)  # populations are column blocks of one matrix
    at, mats = 0, []
    for n, p in zip(n_ind, ploidy):
        whole[:, at : at + n] = rng.binomial(p, f[:, None], size=(n_sites, n))
        mats.append(whole[:, at : at + n])
        at += n
    whole[rng.random(whole.shape) < 0.01] = -1
    mats[2][rng.random(n_sites) < 0.3] = ploidy[2]
    pos = np.cumsum(rng.integers(1, 80, size=n_sites)).astype(np.int32)
    wins = split_genome([int(pos[0]), int(pos[-1])], 50_000, 25_000)
    jobs = [make_job(0, 1, [2], True, u=dict(w=0.4, x=0.1, y_list=[(">=", 0.5)]), q=dict(w=0.4, quantile=0.9, y_list=[(">=", 0.5)])),
            make_job(1, 0, [2], False, u=dict(w=0.6, x=0.0, y_list=[("=", 1.0)]), q=dict(w=0.6, quantile=0.5, y_list=[("=", 1.0)]))]
    more = [make_job(0, 1, [2], False, u=dict(w=0.2, x=0.3, y_list=[("=", 1.0)]), q=dict(w=0.2, quantile=0.95, y_list=[("=", 1.0)]))]
    cap = dict(cap_u=1 << 20, cap_q=1 << 20)
    pg = pack_populations(mats, ploidy, pos)
    want = engine.score(pg, wins, jobs, **cap)
    want_more = engine.score(pg, wins, more, **cap)
    engine.set_host_threads(threads)
    engine.set_i8_wire(wire)
    try:
        got, mg = engine.score_matrices(mats, ploidy, pos, wins, jobs, **cap)
        wire_bytes = engine.i8_wire_bytes()
        got_more = engine.score_resident(more, **cap)
    finally:
        engine.set_host_threads(0)
        engine.set_i8_wire("auto")
    if wire == "dense":
        assert wire_bytes == pg.packed.nbytes
    else:
        assert 0 < wire_bytes <= pg.packed.nbytes + 8 * pg.n_tiles + 64 * pg.n_tiles
    assert [mg.layout.pop[i].bits for i in range(3)] == [pg.layout.pop[i].bits for i in range(3)]
    _same_results(got, want, len(jobs), len(wins))
    _same_results(got_more, want_more, 1, len(wins))
    assert n_sites < 100 or int(want.u.sum()) > 0


@pytest.mark.parametrize("kind", ["incompressible", "all_hom_ref", "sparse"])
def test_int8_pipeline_zt_wire_extremes(kind, engine):
    """The packers' zt records at the extremes: tiles that do not compress (stored raw, flagged in
    the directory), an all-zero matrix (records without payload) and a realistic sparse spectrum
    (wire bytes well under the dense tiles) -- results equal the dense-tile engine's."""
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import make_job
    from sai_b200.windows import split_genome

    rng = np.random.default_rng(len(kind))
    n_sites, n_ind = 40_000, [700, 500, 3]
    if kind == "incompressible":
        mats = [rng.integers(-1, 3, size=(n_sites, n)).astype(np.int8) for n in n_ind]
    elif kind == "all_hom_ref":
        mats = [np.zeros((n_sites, n), dtype=np.int8) for n in n_ind]
    else:
        f = rng.beta(0.15, 3.0, size=n_sites)
        mats = [rng.binomial(2, f[:, None], size=(n_sites, n)).astype(np.int8) for n in n_ind]
        mats[2][rng.random(n_sites) < 0.05] = 2
    pos = np.cumsum(rng.integers(1, 60, size=n_sites)).astype(np.int32)
    wins = split_genome([int(pos[0]), int(pos[-1])], 50_000, 10_000)
    job = make_job(0, 1, [2], False, u=dict(w=0.3, x=0.2, y_list=[(">=", 0.5)]), q=dict(w=0.3, quantile=0.95, y_list=[(">=", 0.5)]))
    pg = pack_populations(mats, [2, 2, 2], pos)
    want = engine.score(pg, wins, [job])
    engine.set_i8_wire("zt")  # whatever record encoder this CPU has
    try:
        got, _ = engine.score_matrices(mats, [2, 2, 2], pos, wins, [job])
        wire_bytes = engine.i8_wire_bytes()
    finally:
        engine.set_i8_wire("auto")
    _same_results(got, want, 1, len(wins))
    if kind == "incompressible":
        assert pg.packed.nbytes <= wire_bytes <= pg.packed.nbytes + 72 * pg.n_tiles
    elif kind == "all_hom_ref":
        assert wire_bytes < 0.05 * pg.packed.nbytes
    else:
        assert wire_bytes < 0.5 * pg.packed.nbytes


def test_int8_pipeline_widens_the_planes_when_the_data_needs_it(engine):
    """Values above the ploidy (the reference's flipped-missing quirk, utils.py:555): the packer
    reports them and score_matrices repeats the call with planes chosen from the data."""
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import make_job
    from sai_b200.windows import split_genome

    rng = np.random.default_rng(5)
    n_sites = 4000
    mats = [rng.integers(0, 3, size=(n_sites, n)).astype(np.int8) for n in (40, 30, 2)]
    mats[1][rng.random(mats[1].shape) < 0.01] = 4  # diploid individual with two flipped missing alleles
    mats[0][rng.random(mats[0].shape) < 0.05] = -1
    pos = np.arange(1, n_sites + 1, dtype=np.int32) * 13
    wins = split_genome([int(pos[0]), int(pos[-1])], 10_000, 5_000)
    job = make_job(0, 1, [2], False, u=dict(w=0.7, x=0.3, y_list=[(">=", 0.5)]), q=dict(w=0.7, quantile=0.9, y_list=[(">=", 0.5)]))
    pg = pack_populations(mats, [2, 2, 2], pos)
    assert pg.layout.pop[1].bits == 3
    want = engine.score(pg, wins, [job])
    got, mg = engine.score_matrices(mats, [2, 2, 2], pos, wins, [job])
    assert mg.layout.pop[1].bits == 3
    _same_results(got, want, 1, len(wins))
    with pytest.raises(ValueError, match="does not fit"):
        engine.score_matrices(mats, [2, 2, 2], pos, wins, [job], bits=[2, 2, 2])
