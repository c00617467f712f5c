"""Pins the CPU oracle (oracle/sai_oracle.py) against
 (a) the known-answer vectors of the reference's own unit tests (transcribed
     with their file:line), and
 (b) fixtures produced by running the reference itself (tests/golden/make_golden.py).
CPU only."""

import json
import os

import numpy as np
import pytest

import sai_oracle as orc
from helpers import GOLDEN, SimplePloidy, SimpleStats, check_items, load_pipe_case, pipe_case_names, vcf_case_names


# ---------------------------------------------------------------- calc_freq KATs
# reference tests/stats/test_stat_utils.py:29-112
@pytest.mark.parametrize(
    "gts, ploidy, expected",
    [
        ([[1, 0, 0, 1], [0, 0, 0, 0], [1, 1, 1, 1]], 1, [0.5, 0.0, 1.0]),
        ([[1, -1, -1, 1], [-1, -1, -1, -1], [1, -1, 1, 1]], 1, [1.0, np.nan, 1.0]),
        ([[1, 1], [0, 0], [2, 2]], 2, [0.5, 0.0, 1.0]),
        ([[1, -1], [0, 0], [-2, 2]], 2, [0.5, 0.0, 1.0]),
        ([[1, 2, 3], [0, 0, 0], [3, 3, 3]], 3, [2 / 3, 0.0, 1.0]),
        ([[2, 2, 2, 2], [1, 3, 0, 4], [0, 0, 0, 0]], 4, [0.5, 0.5, 0.0]),
    ],
)
def test_site_frequency_kat(gts, ploidy, expected):
    got = orc.site_frequency(np.array(gts), ploidy)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(expected))
    np.testing.assert_allclose(got[~np.isnan(got)], np.array(expected)[~np.isnan(expected)], rtol=0, atol=1e-15)


@pytest.mark.parametrize("ploidy", [None, 9.9, -100, 0])
def test_site_frequency_invalid_ploidy(ploidy):  # test_stat_utils.py:100-112
    with pytest.raises(ValueError):
        orc.site_frequency(np.array([[1, 2, 3]]), ploidy)


def test_matching_loci_validation():  # test_stat_utils.py:115-225
    ref = np.array([[0, 1, 0], [1, 1, 0], [0, 0, 1]])
    tgt = np.array([[1, 1, 0], [0, 1, 1], [1, 1, 1]])
    src = [np.array([[0, 0, 1], [1, 1, 0], [0, 1, 1]]), np.array([[1, 1, 0], [1, 0, 0], [1, 1, 0]])]
    for cond in [("=", 0.5), ("<", 0.4), (">", 0.3), ("<=", 0.6), (">=", 0.2)]:
        r, t, c = orc.matching_loci(ref, tgt, src, 0.5, [cond, cond], [2, 2, 2], False)
        assert r.shape == t.shape == c.shape == (3,) and c.dtype == bool
        assert np.all((r >= 0) & (r <= 1)) and np.all((t >= 0) & (t <= 1))
    for w in (-0.1, 1.1):
        with pytest.raises(ValueError, match=r"Parameters w must be within the range \[0, 1\]."):
            orc.matching_loci(ref, tgt, src, w, [("=", 0.5)] * 2, [2, 2, 2], False)
    for y in (-0.1, 1.1):
        with pytest.raises(ValueError, match="Invalid value in y_list"):
            orc.matching_loci(ref, tgt, src, 0.5, [("=", y)], [2, 2, 2], False)
    with pytest.raises(ValueError, match="Invalid operator in y_list"):
        orc.matching_loci(ref, tgt, src, 0.5, [("invalid", 0.5)], [2, 2, 2], False)
    with pytest.raises(ValueError, match="The length of src_gts_list and y_list must match"):
        orc.matching_loci(ref, tgt, src, 0.5, [("=", 0.5)], [2, 2, 2], False)


# ---------------------------------------------------------------- U KATs
# reference tests/stats/test_u_statistic.py:26-209
U_KATS = [
    # (ref, tgt, [src...], ploidies, w, x, y_list, anc, U, positions)
    ([[0, 0, 1], [0, 0, 0], [1, 1, 1]], [[1, 1, 1], [1, 0, 0], [0, 1, 0]], [[[0, 0, 0], [1, 1, 1], [1, 0, 1]]],
     [1, 1, 1], 0.5, 0.5, [("=", 0)], False, 1, [0]),
    ([[0, 0, 1], [0, 0, 0], [1, 1, 1]], [[1, 1, 1], [1, 0, 0], [0, 1, 0]], [[[0, 0, 0], [1, 1, 1], [1, 0, 1]]],
     [1, 1, 1], 0.5, 0.5, [("=", 1)], True, 0, []),
    ([[0, 1, 1], [1, 1, 1]], [[0, 0, 0], [1, 0, 1]], [[[1, 1, 1], [1, 1, 1]]], [1, 1, 1], 0.3, 0.5, [("=", 0)], False, 0, []),
    ([[0, 0, 0], [0, 0, 0]], [[1, 1, 1], [1, 1, 1]], [[[0, 0, 0], [0, 0, 0]]], [1, 1, 1], 0.5, 0.5, [("=", 0)], False, 2, [0, 1]),
    ([[0, 0, 1], [0, 0, 0], [1, 1, 1]], [[0, 1, 1], [0, 0, 1], [1, 1, 1]],
     [[[1, 1, 1], [0, 1, 1], [1, 1, 1]], [[1, 1, 1], [0, 1, 1], [1, 1, 1]]],
     [1, 1, 1, 1], 0.5, 0.5, [("=", 1.0), ("=", 1.0)], False, 1, [0]),
    ([[0, 1, 0], [0, 1, 0], [2, 1, 0]], [[1, 1, 0], [1, 1, 1], [1, 1, 1]], [[[0, 0, 0], [1, 1, 1], [0, 0, 0]]],
     [3, 1, 2], 0.5, 0.5, [("=", 0)], False, 2, [0, 2]),
]


@pytest.mark.parametrize("case", U_KATS)
def test_u_kat(case):
    ref, tgt, srcs, pl, w, x, y_list, anc, exp_u, exp_pos = case
    n = len(ref)
    res = orc.u_statistic(np.array(ref), np.array(tgt), [np.array(s) for s in srcs], pl[0], pl[1], pl[2:],
                          pos=np.arange(n), w=w, x=x, y_list=y_list, anc_allele_available=anc)
    assert res["name"] == "U" and res["value"] == exp_u
    assert np.array_equal(res["cdd_pos"], np.array(exp_pos))


def test_u_missing_kwargs():  # test_u_statistic.py:212-233
    with pytest.raises(ValueError):
        orc.u_statistic(np.zeros((1, 1), int), np.zeros((1, 1), int), [np.zeros((1, 1), int)], 1, 1, [1],
                        pos=np.arange(1), w=0.5, x=0.5, y_list=[("=", 0)])


# ---------------------------------------------------------------- Q KATs
# reference tests/stats/test_q_statistic.py:26-243
Q_KATS = [
    ([[0, 0, 1], [0, 0, 0], [1, 1, 1]], [[0, 1, 1], [0, 0, 1], [1, 1, 1]], [[[1, 1, 1], [0, 1, 1], [1, 1, 1]]],
     [1, 1, 1], 0.5, [("=", 1.0)], 0.95, False, 0.66667, [0]),
    ([[0, 0, 1], [0, 0, 0], [1, 1, 1]], [[0, 1, 1], [0, 0, 1], [1, 1, 1]], [[[1, 1, 1], [0, 1, 1], [1, 1, 1]]],
     [1, 1, 1], 0.5, [("=", 1.0)], 0.95, True, 0.66667, [0]),
    ([[0, 0, 1], [0, 0, 0]], [[0, 1, 1], [1, 1, 1]], [[[1, 1, 1], [1, 1, 1]]], [1, 1, 1], 0.3, [("=", 0.0)], 0.95, False, np.nan, []),
    ([[0, 0, 1], [1, 0, 0], [0, 0, 1]], [[0, 1, 1], [1, 1, 1], [1, 1, 1]], [[[0, 0, 0], [1, 1, 1], [1, 1, 1]]],
     [1, 1, 1], 0.5, [("=", 1.0)], 0.5, False, 1.0, [1, 2]),
    # interpolated edge case, exact expected value 0.9666666666666667 (test_q_statistic.py:148-180)
    ([[0, 0, 1], [0, 0, 0], [1, 1, 1]], [[0, 1, 1], [1, 1, 1], [0, 0, 0]], [[[0, 0, 0], [1, 1, 1], [1, 1, 1]]],
     [1, 1, 1], 0.95, [("=", 1.0)], 0.95, False, 0.9666666666666667, [1]),
    ([[1, 1, 0], [0, 1, 1], [1, 1, 1], [0, 0, 1]], [[0, 0, 0], [1, 1, 1], [1, 1, 1], [1, 1, 1]],
     [[[0, 0, 0], [1, 1, 1], [1, 1, 1], [0, 0, 1]], [[1, 1, 1], [1, 1, 1], [0, 0, 0], [1, 1, 1]]],
     [1, 1, 1, 1], 0.5, [("=", 1), ("=", 1)], 0.95, False, np.nan, []),
    ([[1, 1, 0], [0, 1, 1], [1, 1, 1], [0, 0, 1]], [[0, 0, 0], [1, 1, 1], [1, 1, 1], [1, 1, 1]],
     [[[0, 0, 0], [1, 1, 1], [1, 1, 1], [0, 0, 1]], [[1, 1, 1], [1, 1, 1], [0, 0, 0], [1, 1, 1]]],
     [2, 2, 4, 4], 0.5, [("=", 1), ("=", 1)], 0.95, False, np.nan, []),
]


@pytest.mark.parametrize("case", Q_KATS)
def test_q_kat(case):
    ref, tgt, srcs, pl, w, y_list, q, anc, exp_q, exp_pos = case
    n = len(ref)
    res = orc.q_statistic(np.array(ref), np.array(tgt), [np.array(s) for s in srcs], pl[0], pl[1], pl[2:],
                          pos=np.arange(n), w=w, y_list=y_list, quantile=q, anc_allele_available=anc)
    assert res["name"] == "Q"
    if np.isnan(exp_q):
        assert np.isnan(res["value"])
    else:
        assert np.isclose(res["value"], exp_q)
    assert np.array_equal(res["cdd_pos"], np.array(exp_pos))


def test_q_edge_case_exact():
    """0.9666666666666667 exactly (the other lerp branch gives ...668)."""
    case = Q_KATS[4]
    res = orc.q_statistic(np.array(case[0]), np.array(case[1]), [np.array(case[2][0])], 1, 1, [1], pos=np.arange(3),
                          w=0.95, y_list=[("=", 1.0)], quantile=0.95, anc_allele_available=False)
    assert float(res["value"]) == 0.9666666666666667


def test_q_missing_kwargs():  # test_q_statistic.py:246-268
    with pytest.raises(ValueError):
        orc.q_statistic(np.zeros((1, 1), int), np.zeros((1, 1), int), [np.zeros((1, 1), int)], 1, 1, [1],
                        pos=np.arange(1), w=0.5, quantile=0.95, anc_allele_available=False)


# ---------------------------------------------------------------- geometry goldens
def test_split_genome_goldens():  # tests/utils/test_utils.py:423-447
    pos = np.array([0, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100])
    assert orc.split_genome(pos, 30, 20) == [(1, 30), (21, 50), (41, 70), (61, 90), (81, 110)]
    with pytest.raises(ValueError, match="`step_size` cannot be greater than `window_size`"):
        orc.split_genome(np.array([0, 10, 20]), 20, 25)
    with pytest.raises(ValueError, match="`pos` array must not be empty"):
        orc.split_genome(np.array([]), 30, 10)


def test_chunk_ranges_golden():  # tests/generators/test_chunk_generator.py:25-40 (chr21 of test.data.vcf: 2309..48989)
    wins = orc.split_genome([2309, 48989], 10000, 5000)
    assert orc.split_windows_ranges(wins, 2) == [(1, 30000), (25001, 55000)]
    # a worker re-derives exactly its own windows from its range (window_generator.py:132-144)
    assert orc.chunk_windows(1, 30000, 10000, 5000) == wins[:5]
    assert orc.chunk_windows(25001, 55000, 10000, 5000) == wins[5:]
    assert orc.chunk_windows(0, 6666, 6666, 6666) == [(1, 6666)]


# ---------------------------------------------------------------- quantile restatement
def test_quantile_linear_matches_numpy():
    rng = np.random.default_rng(7)
    for _ in range(3000):
        n = int(rng.integers(1, 40))
        d = int(rng.integers(1, 60))
        v = rng.integers(0, d + 1, size=n) / d
        if rng.random() < 0.3:
            v = 1 - v
        q = float(rng.choice([0, 0.25, 0.5, 0.9, 0.95, 0.99, 1.0, rng.random()]))
        assert orc.quantile_linear(v, q) == float(np.nanquantile(v, q))


# ---------------------------------------------------------------- reference-generated fixtures
def test_dd_kat():  # tests/stats/test_dd_statistic.py:25-46
    ref_gts = np.array([[1, 1], [0, 0]])
    tgt_gts = np.array([[1, 0], [0, 1]])
    src_gts = np.array([[0, 1], [1, 1]])
    assert np.isclose(orc.dd_statistic(ref_gts, tgt_gts, [src_gts])[0], 0.5)


def test_dd_cases_against_reference_outputs():
    """DD of the 400 random cases (missing calls as -1 / -2, ploidy 1-4) as the
    reference's DdStatistic (scipy cdist) returned it: bit-exact."""
    meta = json.load(open(os.path.join(GOLDEN, "stat_cases.json")))
    arrs = np.load(os.path.join(GOLDEN, "stat_cases.npz"))
    for c, m in enumerate(meta):
        mats = [arrs[f"c{c}_g{k}"].astype(np.int64) for k in range(2 + m["n_src"])]
        got = orc.dd_statistic(mats[0], mats[1], mats[2:])
        assert [float(v).hex() for v in got] == m["DD"], c


def test_stat_cases_against_reference_outputs():
    meta = json.load(open(os.path.join(GOLDEN, "stat_cases.json")))
    arrs = np.load(os.path.join(GOLDEN, "stat_cases.npz"))
    for c, m in enumerate(meta):
        mats = [arrs[f"c{c}_g{k}"].astype(np.int64) for k in range(2 + m["n_src"])]
        pos = arrs[f"c{c}_pos"]
        y_list = [tuple(y) for y in m["y_list"]]
        pl = m["ploidy"]
        ru = orc.u_statistic(mats[0], mats[1], mats[2:], pl[0], pl[1], pl[2:], pos=pos, w=m["w"], x=m["x"],
                             y_list=y_list, anc_allele_available=m["anc"])
        rq = orc.q_statistic(mats[0], mats[1], mats[2:], pl[0], pl[1], pl[2:], pos=pos, w=m["w"], quantile=m["q"],
                             y_list=y_list, anc_allele_available=m["anc"])
        assert ru["value"] == m["U"], c
        assert [int(p) for p in ru["cdd_pos"]] == m["U_pos"], c
        if m["Q"] == "nan":
            assert np.isnan(rq["value"]), c
        else:
            assert float(rq["value"]).hex() == m["Q"], c
        assert [int(p) for p in rq["cdd_pos"]] == m["Q_pos"], c


@pytest.mark.parametrize("name", pipe_case_names())
def test_pipeline_against_reference_outputs(name):
    case, pos, data = load_pipe_case(name)
    mk = lambda d: {p: orc.PopData(pos, m.astype(np.int64)) for p, m in d.items()}
    stats = SimpleStats(case["stats"])
    items = orc.score_chunk(case["chr_name"], case["start"], case["end"], case["win_len"], case["win_step"],
                            mk(data["ref"]), mk(data["tgt"]), mk(data["src"]), SimplePloidy(case["ploidies"]),
                            stats, case["anc"], out_data=mk(data["outgroup"]) if "outgroup" in data else None)
    check_items(items, case["items"])
    rows, logs = orc.format_items(items, [s for s in case["stats"] if s in ("U", "Q") or case["stats"][s] is True])
    assert "".join(rows) == case["text"]["tsv"]
    for key in logs:
        assert "".join(logs[key]) == case["text"][key]


@pytest.mark.parametrize("name", vcf_case_names())
def test_vcf_fixtures_against_reference_outputs(name):
    """Host ingest (sai_b200.vcf) + oracle reproduce the reference pipeline
    goldens: Q == 0.9 (tests/test_sai.py:63), U == 3
    (tests/preprocessors/test_feature_preprocessor.py:223), U rows [0, 1]
    (tests/test_sai.py:150-151)."""
    from sai_b200 import vcf as V

    case = json.load(open(os.path.join(GOLDEN, f"vcf_{name}.json")))
    pc = SimplePloidy(case["ploidies"])
    anc = os.path.join(GOLDEN, f"vcf_{name}.anc.bed") if case["anc"] else None
    out_list = os.path.join(GOLDEN, f"vcf_{name}.outgroup.list")
    groups = V.read_data(os.path.join(GOLDEN, case["vcf"]), case["chr_name"], pc,
                         *[os.path.join(GOLDEN, f"vcf_{name}.{g}.list") for g in ("ref", "tgt", "src")],
                         out_list if os.path.exists(out_list) else None, anc, start=case["start"], end=case["end"])
    mk = lambda d: {p: orc.PopData(x.POS, x.GT.astype(np.int64)) for p, x in d.items()}
    items = orc.score_chunk(case["chr_name"], case["start"], case["end"], case["win_len"], case["win_step"],
                            mk(groups["ref"][0]), mk(groups["tgt"][0]), mk(groups["src"][0]), pc,
                            SimpleStats(case["stats"]), case["anc"],
                            out_data=mk(groups["outgroup"][0]) if groups["outgroup"][0] is not None else None)
    check_items(items, case["items"])
    if name == "outgroup_stats":  # the reference's stored golden, tests/data/test.with.outgroup.res.tsv (test_sai.py:92-110)
        want = dict(fd=0.0012826844929596443, df=0.0012417913767941238, Danc=-0.12498082112760132, Dplus=-0.12149240420484364)
        for k, v in want.items():
            assert np.isclose(items[0][k][0], v) and items[0][k][0] == v
    if name == "example_q":
        assert float(items[0]["Q"]) == 0.9
    if name == "example_u":
        assert items[0]["U"] == 3
    if name == "mixed_ploidy":
        assert [it["U"] for it in items] == [0, 1]


def test_mixed_ploidy_df_kat():
    """The reference's KAT for df with two sources on the mixed-ploidy VCF
    (tests/test_sai.py:127-151: df.src1 row 0 == -0.6086956521739131, df.src2 row 1 == -0.45454545454545453,
    U rows [0, 1], config df: True / fd: False) through the host ingest + the oracle."""
    from sai_b200 import vcf as V

    name = "mixed_ploidy"
    case = json.load(open(os.path.join(GOLDEN, f"vcf_{name}.json")))
    pc = SimplePloidy(case["ploidies"])
    stats = SimpleStats({"df": True, "fd": False, "U": case["stats"]["U"]})
    groups = V.read_data(os.path.join(GOLDEN, case["vcf"]), case["chr_name"], pc,
                         *[os.path.join(GOLDEN, f"vcf_{name}.{g}.list") for g in ("ref", "tgt", "src")], None,
                         os.path.join(GOLDEN, f"vcf_{name}.anc.bed"), start=case["start"], end=case["end"])
    mk = lambda d: {p: orc.PopData(x.POS, x.GT.astype(np.int64)) for p, x in d.items()}
    items = orc.score_chunk(case["chr_name"], case["start"], case["end"], case["win_len"], case["win_step"],
                            mk(groups["ref"][0]), mk(groups["tgt"][0]), mk(groups["src"][0]), pc, stats, True)
    assert [it["U"] for it in items] == [0, 1] and all("fd" not in it for it in items)
    assert np.isclose(items[0]["df"][0], -0.6086956521739131) and np.isclose(items[1]["df"][1], -0.45454545454545453)


def test_pairwise_sum_model_matches_numpy():
    """The summation order the pattern kernel reproduces IS numpy's: np.sum of a contiguous float64
    vector against the spelled-out pairwise model, for every structural case (n < 8, one block,
    tails, recursion, beyond the 8192-element iterator buffer)."""
    rng = np.random.default_rng(3)
    sizes = list(range(0, 40)) + [63, 64, 65, 127, 128, 129, 130, 200, 255, 256, 257, 1000, 1199, 1200, 1201,
                                  4096, 8191, 8192, 8193, 8200, 20000, 50001]
    for n in sizes:
        for rep in range(3):
            a = (rng.random(n) ** 3) * rng.choice([1e-3, 1.0, 1e3], size=n)
            assert orc.pairwise_sum_model(a) == float(np.sum(a)), (n, rep)


# ---------------------------------------------------------------- the installed reference (oracle/_ref)
def test_installed_reference_pool_matches_the_oracle():
    """oracle/_ref (the unmodified reference, pip-installed by oracle/build_ref.py) driven through its
    own mp_pool + _split_windows_ranges -- the CPU arm of bench.py -- gives the oracle's items, with
    and without worker processes (window-range chunks with their halo == one chunk)."""
    import ref_driver

    if not ref_driver.available():
        pytest.skip("oracle/_ref not installed (python oracle/build_ref.py where /root/reference is mounted)")
    import synth
    from sai_b200.configs import PloidyConfig, StatConfig

    pops = {"ref": {"AFR": (90, 2)}, "tgt": {"EUR": (70, 2)}, "src": {"NEA": (2, 2), "DEN": (1, 2)}}
    pos, mats = synth.make_populations(11, 5000, pops, mean_gap=60.0, introgressed=0.03, missing=0.02, src_all_missing=0.003)
    pl = {g: {p: v[1] for p, v in pops[g].items()} for g in pops}
    stats = {"U": {"ref": {"AFR": 0.05}, "tgt": {"EUR": 0.3}, "src": {"NEA": "=1", "DEN": ">=0.5"}},
             "Q": {"ref": {"AFR": 0.05}, "tgt": {"EUR": 0.9}, "src": {"NEA": "=1", "DEN": ">=0.5"}}}
    i64 = lambda d: {p: m.astype(np.int64) for p, m in d.items()}
    ref_driver.set_data("t", pos, i64(mats["ref"]), i64(mats["tgt"]), i64(mats["src"]))
    _, _, ref_split = ref_driver.make_classes()
    wins = ref_split([int(pos[0]), int(pos[-1])], 20000, 5000)
    assert wins == orc.split_genome([int(pos[0]), int(pos[-1])], 20000, 5000)
    pooled = ref_driver.score_windows("t", "1", wins, 7, 2, 20000, 5000, pl, stats, False)
    serial = ref_driver.score_windows("t", "1", wins, 1, 1, 20000, 5000, pl, stats, False)
    mk = lambda d: {p: orc.PopData(pos, m.astype(np.int64)) for p, m in d.items()}
    exp = orc.score_chunk("1", wins[0][0], wins[-1][1], 20000, 5000, mk(mats["ref"]), mk(mats["tgt"]), mk(mats["src"]),
                          PloidyConfig(pl), StatConfig(stats), False)
    assert len(pooled) == len(serial) == len(exp) == len(wins)
    same = lambda a, b: (a != a and b != b) or a == b
    for a, b, e in zip(pooled, serial, exp):
        for o in (b, e):
            assert (a["start"], a["end"], a["nsnps"], a["src_pop_list"]) == (o["start"], o["end"], o["nsnps"], tuple(o["src_pop_list"]))
            assert same(a["U"], o["U"]) and same(a["Q"], o["Q"])
            assert np.array_equal(a["cdd_pos"]["U"], o["cdd_pos"]["U"]) and np.array_equal(a["cdd_pos"]["Q"], o["cdd_pos"]["Q"])
    assert sum(a["U"] for a in pooled if a["U"] == a["U"]) > 0
