"""Edge cases of the CUDA path (through the C ABI): empty and degenerate inputs,
extreme coordinates, the capacity limits of the ABI, parameter extremes, and a
hypothesis-driven differential test against the oracle."""

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import sai_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from sai_b200.scoring import HostEngine

    e = HostEngine(0)
    yield e
    e.close()


def _score(engine, mats, ploidy, pos, wins, n_src, anc, u=None, q=None):
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import make_job

    pg = pack_populations(mats, ploidy, pos)
    job = make_job(0, 1, list(range(2, 2 + n_src)), anc, u, q)
    return engine.score(pg, wins, [job])


def _expect(mats, ploidy, pos, win, n_src, anc, u, q):
    keep = (pos >= win[0]) & (pos <= win[1])
    sub = [m[keep].astype(np.int64) for m in mats]
    if not keep.any():
        return 0, None, None
    eu = orc.u_statistic(sub[0], sub[1], sub[2 : 2 + n_src], ploidy[0], ploidy[1], list(ploidy[2 : 2 + n_src]),
                         pos=pos[keep], anc_allele_available=anc, **u)
    eq = orc.q_statistic(sub[0], sub[1], sub[2 : 2 + n_src], ploidy[0], ploidy[1], list(ploidy[2 : 2 + n_src]),
                         pos=pos[keep], anc_allele_available=anc, **q)
    return int(keep.sum()), eu, eq


def _compare(res, i, exp):
    n, eu, eq = exp
    assert res.nsnps[0, i] == n
    if eu is None:
        assert res.u[0, i] == 0 and np.isnan(res.q[0, i]) and res.q_cnt[0, i] == 0
        return
    assert res.u[0, i] == eu["value"]
    assert np.array_equal(res.u_positions(0, i), eu["cdd_pos"])
    if np.isnan(eq["value"]):
        assert np.isnan(res.q[0, i])
    else:
        assert res.q[0, i] == float(eq["value"])
    assert np.array_equal(res.q_positions(0, i), np.asarray(eq["cdd_pos"], dtype=np.int32))


U = dict(w=0.5, x=0.2, y_list=[("=", 1.0)])
Q = dict(w=0.5, quantile=0.95, y_list=[("=", 1.0)])


def test_empty_inputs(engine):
    z = [np.zeros((0, 3), np.int8), np.zeros((0, 2), np.int8), np.zeros((0, 1), np.int8)]
    res = _score(engine, z, [2, 2, 2], np.zeros(0, np.int32), [(1, 100), (50, 60)], 1, True, U, Q)
    assert res.nsnps.tolist() == [[0, 0]] and res.u.tolist() == [[0, 0]] and np.isnan(res.q).all()
    rng = np.random.default_rng(0)
    m = [rng.integers(0, 3, size=(5, 3)).astype(np.int8) for _ in range(3)]
    res = _score(engine, m, [2, 2, 2], np.arange(1, 6), [], 1, True, U, Q)
    assert res.nsnps.shape == (1, 0) and res.totals.tolist() == [[0, 0]]


def test_degenerate_windows_and_coordinates(engine):
    rng = np.random.default_rng(1)
    n = 300
    top = 2_147_483_000
    pos = np.sort(rng.choice(np.arange(top - 100_000, top), size=n, replace=False)).astype(np.int32)
    mats = [rng.integers(0, 2, size=(n, 12)).astype(np.int8), rng.integers(0, 3, size=(n, 9)).astype(np.int8),
            np.where(rng.random((n, 2)) < 0.5, 2, 0).astype(np.int8)]
    wins = [(1, 10), (int(pos[0]), int(pos[0])), (int(pos[-1]), int(pos[-1]) + 5), (int(pos[-1]) + 1, 2**40),
            (1, 2**40), (int(pos[10]), int(pos[200])), (int(pos[10]), int(pos[200])), (int(pos[50]) + 1, int(pos[51]) - 1),
            (2**33, 2**34), (int(pos[100]), int(pos[100]) + 3), (int(pos[200]), int(pos[10])), (2**40, 1)]
    res = _score(engine, mats, [2, 2, 2], pos, wins, 1, False, U, Q)
    for i, w in enumerate(wins):
        _compare(res, i, _expect(mats, [2, 2, 2], pos, w, 1, False, U, Q))
    assert res.nsnps[0, 4] == n and res.nsnps[0, 0] == 0 and res.nsnps[0, 8] == 0
    assert res.nsnps[0, 10] == 0 and res.nsnps[0, 11] == 0  # end < start: empty, like the reference's masks


def test_many_windows_in_any_order(engine):
    """More windows than resident warps, so that every warp walks a run of windows
    and reuses the previous search bounds as hints: sorted, reversed, shuffled
    and repeated windows must all give the result of the window on its own."""
    rng = np.random.default_rng(5)
    n = 30_000
    pos = np.cumsum(rng.integers(1, 60, size=n)).astype(np.int32)
    mats = [rng.integers(0, 2, size=(n, 20)).astype(np.int8), rng.integers(0, 3, size=(n, 14)).astype(np.int8),
            np.where(rng.random((n, 2)) < 0.6, 2, 0).astype(np.int8)]
    top = int(pos[-1])
    starts = np.sort(rng.integers(1, top, size=600))
    base = [(int(s), int(s) + int(l)) for s, l in zip(starts, rng.integers(0, 40_000, size=600))]
    base += [(1, 2**40), (top + 1, top + 10), (1, 1)]
    ref = _score(engine, mats, [2, 2, 2], pos, base, 1, True, U, Q)  # one window per warp: no hints carried
    for i in rng.choice(len(base), size=40, replace=False):
        _compare(ref, int(i), _expect(mats, [2, 2, 2], pos, base[int(i)], 1, True, U, Q))
    order = np.concatenate([
        np.sort(rng.integers(0, len(base), size=30_000)),          # sorted by start (hints hold)
        np.sort(rng.integers(0, len(base), size=15_000))[::-1],    # descending (every hint must be dropped)
        rng.integers(0, len(base), size=25_000),                   # shuffled
    ])
    wins = [base[i] for i in order]
    res = _score(engine, mats, [2, 2, 2], pos, wins, 1, True, U, Q)
    assert np.array_equal(res.nsnps[0], ref.nsnps[0][order])
    assert np.array_equal(res.u[0], ref.u[0][order])
    assert np.array_equal(res.q[0], ref.q[0][order], equal_nan=True)
    assert np.array_equal(res.q_cnt[0], ref.q_cnt[0][order])
    for i in rng.choice(len(wins), size=300, replace=False):
        assert np.array_equal(res.u_positions(0, int(i)), ref.u_positions(0, int(order[i])))
        assert np.array_equal(res.q_positions(0, int(i)), ref.q_positions(0, int(order[i])))


def test_abi_limits(engine):
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import make_job

    rng = np.random.default_rng(2)
    n = 500
    pos = np.arange(10, 10 + 7 * n, 7)
    # 8 sources (SAI_MAX_SRC), 16 populations (SAI_MAX_POPS) in one matrix, 8 jobs (SAI_MAX_JOBS) in one pass
    ploidy = [2, 2] + [1, 2, 1, 2, 4, 1, 2, 8] + [2] * 6
    mats = [rng.integers(0, 2, size=(n, 40)).astype(np.int8), rng.integers(0, 3, size=(n, 33)).astype(np.int8)]
    for p in ploidy[2:]:
        g = np.where(rng.random((n, 2)) < 0.7, p, rng.integers(0, p + 1, size=(n, 2))).astype(np.int8)
        g[rng.random((n, 2)) < 0.02] = -1
        mats.append(g)
    assert len(mats) == 16
    pg = pack_populations(mats, ploidy, pos)
    assert pg.layout.pop[9].bits == 4 and pg.layout.pop[6].bits == 3
    wins = [(1, 1200), (600, 2400), (1, 10**6)]
    jobs, specs = [], []
    for j in range(8):
        k = 8 - j  # job j uses the first k sources
        ys = [(">=", 0.5)] * k
        u, q = dict(w=0.6, x=0.1 * j, y_list=ys), dict(w=0.6, quantile=0.1 * j, y_list=ys)
        jobs.append(make_job(0, 1, list(range(2, 2 + k)), j % 2 == 0, u, q))
        specs.append((k, j % 2 == 0, u, q))
    res = engine.score(pg, wins, jobs)
    for j, (k, anc, u, q) in enumerate(specs):
        for i, w in enumerate(wins):
            n_exp, eu, eq = _expect(mats, ploidy, pos, w, k, anc, u, q)
            assert res.nsnps[j, i] == n_exp and res.u[j, i] == eu["value"]
            assert np.array_equal(res.u_positions(j, i), eu["cdd_pos"])
            assert (np.isnan(res.q[j, i]) and np.isnan(eq["value"])) or res.q[j, i] == float(eq["value"])
            assert np.array_equal(res.q_positions(j, i), np.asarray(eq["cdd_pos"], dtype=np.int32))
    assert res.u.sum() > 0
    with pytest.raises(ValueError, match="at most 8 source populations"):
        make_job(0, 1, list(range(2, 11)), True, dict(w=0.1, x=0.1, y_list=[("=", 1.0)] * 9))
    with pytest.raises(ValueError, match="between 1 and 8 jobs"):
        engine.score(pg, wins, jobs + jobs[:1])
    with pytest.raises(ValueError, match="at most 16 populations"):
        pack_populations(mats + mats[:1], ploidy + [2], pos)


@pytest.mark.parametrize("u, q", [
    (dict(w=0.0, x=0.0, y_list=[("=", 1.0)]), dict(w=0.0, quantile=0.5, y_list=[("=", 1.0)])),   # nothing is < 0
    (dict(w=1.0, x=1.0, y_list=[(">=", 0.0)]), dict(w=1.0, quantile=0.0, y_list=[(">=", 0.0)])),  # nothing is > 1; q = min
    (dict(w=1.0, x=0.0, y_list=[("<=", 1.0)]), dict(w=1.0, quantile=1.0, y_list=[("<=", 1.0)])),  # q = max
    (dict(w=0.3, x=0.5, y_list=[("<", 0.5)]), dict(w=0.3, quantile=0.25, y_list=[(">", 0.5)])),
    (dict(w=0.3, x=0.5, y_list=[("=", 0.5)]), dict(w=0.3, quantile=0.75, y_list=[("=", 0.5)])),  # y = 1 - y: every match inverts
])
@pytest.mark.parametrize("anc", [True, False])
def test_parameter_extremes(engine, u, q, anc):
    rng = np.random.default_rng(3)
    n = 2000
    pos = np.cumsum(rng.integers(1, 30, size=n)).astype(np.int32)
    f = rng.beta(0.5, 0.5, size=n)
    mats = [rng.binomial(2, f[:, None] * 0.3, size=(n, 25)).astype(np.int8), rng.binomial(2, f[:, None], size=(n, 17)).astype(np.int8),
            rng.binomial(2, np.round(f)[:, None] * 0.5 + 0.25 * (rng.random((n, 1)) < 0.3), size=(n, 2)).astype(np.int8)]
    for m in mats:
        m[rng.random(m.shape) < 0.03] = -1
    wins = [(1, 5000), (2000, 9000), (1, int(pos[-1]))]
    res = _score(engine, mats, [2, 2, 2], pos, wins, 1, anc, u, q)
    for i, w in enumerate(wins):
        _compare(res, i, _expect(mats, [2, 2, 2], pos, w, 1, anc, u, q))


gt_matrix = lambda n_sites, n_ind, ploidy: st.lists(
    st.lists(st.integers(-2, ploidy), min_size=n_ind, max_size=n_ind), min_size=n_sites, max_size=n_sites)


@st.composite
def stat_case(draw):
    n_sites = draw(st.integers(1, 40))
    n_src = draw(st.integers(1, 3))
    ploidy = [draw(st.integers(1, 4)) for _ in range(2 + n_src)]
    n_ind = [draw(st.integers(1, 5)) for _ in range(2 + n_src)]
    mats = [np.array(draw(gt_matrix(n_sites, n, p)), dtype=np.int8) for n, p in zip(n_ind, ploidy)]
    ops = st.sampled_from(["=", "<", ">", "<=", ">="])
    ys = st.sampled_from([0.0, 0.25, 1 / 3, 0.5, 2 / 3, 0.8, 1.0])
    y_list = [(draw(ops), draw(ys)) for _ in range(n_src)]
    w = draw(st.sampled_from([0.0, 0.1, 0.3, 0.5, 1.0]))
    x = draw(st.sampled_from([0.0, 0.2, 0.5, 0.9]))
    q = draw(st.sampled_from([0.0, 0.1, 0.5, 0.9, 0.95, 1.0]))
    return mats, ploidy, n_src, y_list, w, x, q, draw(st.booleans())


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(case=stat_case())
def test_hypothesis_differential(engine, case):
    mats, ploidy, n_src, y_list, w, x, q, anc = case
    n = mats[0].shape[0]
    pos = np.arange(1, n + 1, dtype=np.int32) * 3
    u, qq = dict(w=w, x=x, y_list=y_list), dict(w=w, quantile=q, y_list=y_list)
    wins = [(1, 3 * n), (1, 3 * (n // 2) + 1)]
    res = _score(engine, mats, ploidy, pos, wins, n_src, anc, u, qq)
    for i, win in enumerate(wins):
        _compare(res, i, _expect(mats, ploidy, pos, win, n_src, anc, u, qq))


@pytest.mark.parametrize("shape", ["sparse", "dense", "mixed_bits", "config2_like"])
def test_zt_wire_format_gpu(engine, shape):
    """The zero-suppressed wire format: the device decoder rebuilds the packed tiles bit for bit,
    and the engine gives identical results from the zt stream and from the dense tiles."""
    import ctypes as C

    import torch

    from sai_b200 import _cabi
    from sai_b200.encode import compress, pack_populations
    from sai_b200.scoring import make_job

    rng = np.random.default_rng(31)
    if shape == "sparse":
        n, sizes, ploidy = 5000, (300, 90, 4), [2, 2, 2]
        f = rng.beta(0.2, 2.0, size=n)
        mats = [rng.binomial(2, f[:, None], size=(n, k)).astype(np.int8) for k in sizes]
        mats[2][:] = np.where(rng.random((n, 1)) < 0.3, 2, mats[2])
        mats[1][rng.random(mats[1].shape) < 0.002] = -1
    elif shape == "dense":
        n, sizes, ploidy = 700, (64, 96, 32), [2, 2, 2]
        mats = [rng.integers(-1, 3, size=(n, k)).astype(np.int8) for k in sizes]
    elif shape == "mixed_bits":
        n, ploidy, sizes = 2049, [4, 3, 8], (45, 33, 7)
        f = rng.beta(0.3, 3.0, size=n)
        mats = [rng.binomial(p, f[:, None], size=(n, k)).astype(np.int8) for k, p in zip(sizes, ploidy)]
        mats[0][rng.random(mats[0].shape) < 0.01] = -2
    else:
        n, sizes, ploidy = 40_000, (1500, 1000, 4), [2, 2, 2]
        f = rng.random(n) ** 4
        mats = [rng.binomial(2, f[:, None], size=(n, k)).astype(np.int8) for k in sizes]
        mats[2][:] = np.where(rng.random((n, 1)) < 0.01, 2, mats[2])
    pos = np.cumsum(rng.integers(1, 80, size=n)).astype(np.int32)
    pg = pack_populations(mats, ploidy, pos)
    zt = compress(pg)
    # device decoder vs the packed tiles, in two tile ranges
    lib = _cabi.load()
    d_stream = torch.from_numpy(zt.stream.copy()).cuda()
    d_off = torch.from_numpy(zt.tile_off.view(np.int64).copy()).cuda()
    d_packed = torch.full((pg.packed.nbytes,), 0xAB, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    half = pg.n_tiles // 2
    for t0, nt in ((half, pg.n_tiles - half), (0, half)):
        _cabi.check(lib.sai_zt_decode(C.byref(pg.layout), d_stream.data_ptr(), d_stream.numel(), d_off.data_ptr(), t0, nt,
                                      d_packed.data_ptr(), st))
    assert np.array_equal(d_packed.cpu().numpy(), pg.packed)
    # engine: zt stream vs dense tiles
    top = int(pos[-1])
    wins = [(s, s + 49_999) for s in range(1, top, 10_000)]
    u = dict(w=0.3, x=0.1, y_list=[(">=", 0.5)])
    q = dict(w=0.3, quantile=0.95, y_list=[(">=", 0.5)])
    job = make_job(0, 1, [2], True, u, q)
    a, b = engine.score(pg, wins, [job]), engine.score(zt, wins, [job])
    for name in ("nsnps", "u", "q_cnt"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert np.array_equal(a.q, b.q, equal_nan=True) and a.nsnps.sum() > 0
    if shape in ("sparse", "config2_like"):
        assert a.u.sum() > 0 and np.isfinite(a.q).any()
    for i in range(len(wins)):
        assert np.array_equal(a.u_positions(0, i), b.u_positions(0, i))
        assert np.array_equal(a.q_positions(0, i), b.q_positions(0, i))
    assert (zt.stream.nbytes < pg.packed.nbytes) == (shape != "dense")


def test_integer_fast_path_thresholds(engine):
    """Fully called sites are decided by integer interval tests whose ends the host finds by
    bisection (site_cond.cuh); sites with a missing call take the float64 division path.  Every
    count combination of a small layout x thresholds sitting exactly on, just below and just above
    every attainable frequency (and 1 - frequency) x every operator, with and without ancestral
    alleles, with and without one missing call per site: all against the oracle."""
    from sai_b200.encode import pack_populations
    from sai_b200.scoring import make_job

    rng = np.random.default_rng(77)
    n_ref, n_tgt, n_src = 3, 2, 1  # diploid: denominators 6, 4, 2
    combos = [(a, b, c) for a in range(7) for b in range(5) for c in range(3)]

    def column(total, n_ind):
        g = np.zeros(n_ind, np.int8)
        for i in range(n_ind):
            g[i] = min(2, total)
            total -= g[i]
        return g

    base = [np.array([column(c[k], n) for c in combos], dtype=np.int8) for k, n in enumerate((n_ref, n_tgt, n_src))]
    # second half: the same sites with one extra individual per population that is missing -> division path
    with_missing = [np.concatenate([m, np.full((len(combos), 1), -1, np.int8)], axis=1) for m in base]
    crit = sorted({n / d for d in (6, 4, 2, 8, 3) for n in range(d + 1)} | {1 - n / d for d in (6, 4, 2) for n in range(d + 1)})
    crit = sorted({float(v) for c in crit for v in (c, np.nextafter(c, 2.0), np.nextafter(c, -1.0)) if 0.0 <= v <= 1.0})
    ops = ["=", "<", ">", "<=", ">="]
    pos = np.arange(1, len(combos) + 1, dtype=np.int32)
    wins = [(int(p), int(p)) for p in pos]
    checked = 0
    for mats in (base, with_missing):
        pg = pack_populations(mats, [2, 2, 2], pos)
        m64 = [m.astype(np.int64) for m in mats]
        for batch in range(6):
            specs, jobs = [], []
            for _ in range(8):
                w, x, y, q = (float(rng.choice(crit)) for _ in range(4))
                op, anc = str(rng.choice(ops)), bool(rng.integers(0, 2))
                yq, opq = (y, op) if rng.random() < 0.5 else (float(rng.choice(crit)), str(rng.choice(ops)))
                u = dict(w=w, x=x, y_list=[(op, y)])
                qd = dict(w=w if rng.random() < 0.5 else float(rng.choice(crit)), quantile=q, y_list=[(opq, yq)])
                specs.append((anc, u, qd))
                jobs.append(make_job(0, 1, [2], anc, u, qd))
            res = engine.score(pg, wins, jobs)
            for j, (anc, u, qd) in enumerate(specs):
                for i in range(len(combos)):
                    sub = [m[i : i + 1] for m in m64]
                    eu = orc.u_statistic(sub[0], sub[1], [sub[2]], 2, 2, [2], pos=pos[i : i + 1], anc_allele_available=anc, **u)
                    eq = orc.q_statistic(sub[0], sub[1], [sub[2]], 2, 2, [2], pos=pos[i : i + 1], anc_allele_available=anc, **qd)
                    assert res.u[j, i] == eu["value"], (j, i, u, anc, combos[i])
                    if np.isnan(eq["value"]):
                        assert np.isnan(res.q[j, i]), (j, i, qd, anc, combos[i])
                    else:
                        assert res.q[j, i] == float(eq["value"]), (j, i, qd, anc, combos[i])
                    checked += 1
    assert checked == 2 * 6 * 8 * len(combos)


@pytest.mark.gpu
def test_multiallelic_values_above_14_score_like_the_oracle():
    """Multi-allelic genotype indices are not recoded by the reference (alt_number=1 only limits
    ALT, utils.py:119-135), so a diploid sum can exceed 14: the layout widens to 5..8 planes and
    U / Q / candidate lists stay identical to the oracle (most such sites are simply invalid:
    frequency above 1)."""
    from sai_b200.scoring import HostEngine, make_job
    from sai_b200.windows import split_genome

    rng = np.random.default_rng(77)
    n_sites = 3000
    f = rng.beta(0.3, 1.5, size=n_sites)
    mats = [rng.binomial(2, f[:, None], size=(n_sites, n)).astype(np.int8) for n in (60, 45, 3)]
    mats[2][rng.random(n_sites) < 0.3] = 2
    for m, vmax in zip(mats, (16, 40, 127)):  # a few multi-allelic calls per population
        hit = rng.random(m.shape) < 0.003
        m[hit] = rng.integers(3, vmax + 1, size=int(hit.sum()))
    mats[0][rng.random(mats[0].shape) < 0.02] = -1
    pos = np.cumsum(rng.integers(1, 60, size=n_sites)).astype(np.int32)
    wins = split_genome([int(pos[0]), int(pos[-1])], 20_000, 5_000)
    u_kw = dict(w=0.5, x=0.2, y_list=[(">=", 0.5)])
    q_kw = dict(w=0.5, quantile=0.9, y_list=[(">=", 0.5)])
    eng = HostEngine(0)
    got, mg = eng.score_matrices(mats, [2, 2, 2], pos, wins, [make_job(0, 1, [2], False, u=u_kw, q=q_kw)])
    eng.close()
    assert [mg.layout.pop[i].bits for i in range(3)] == [5, 6, 8]
    m64 = [m.astype(np.int64) for m in mats]
    for i, (s, e) in enumerate(wins):
        keep = (pos >= s) & (pos <= e)
        eu = orc.u_statistic(m64[0][keep], m64[1][keep], [m64[2][keep]], 2, 2, [2], pos=pos[keep], anc_allele_available=False, **u_kw)
        eq = orc.q_statistic(m64[0][keep], m64[1][keep], [m64[2][keep]], 2, 2, [2], pos=pos[keep], anc_allele_available=False, **q_kw)
        assert got.nsnps[0, i] == keep.sum() and got.u[0, i] == eu["value"]
        assert np.array_equal(got.u_positions(0, i), eu["cdd_pos"])
        assert (np.isnan(got.q[0, i]) and np.isnan(eq["value"])) or got.q[0, i] == float(eq["value"])
        assert np.array_equal(got.q_positions(0, i), np.asarray(eq["cdd_pos"], dtype=np.int32))
    assert int(got.u.sum()) > 0
