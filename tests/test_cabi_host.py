"""CPU-only checks of the C-ABI library and the host logic (no compute calls)."""

import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "sai_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sai_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sai_b200 import _cabi

    lib = _cabi.load()
    declared = _declared_functions()
    assert len(declared) >= 17
    assert sorted(_cabi.SYMBOLS) == declared  # the ctypes table mirrors the header
    for name in declared:
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_cabi.lib_path())], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", nm), name
    assert b"sm_100a" in lib.sai_version()


def test_struct_sizes_match_header():
    from sai_b200 import _cabi

    assert C.sizeof(_cabi.PopLayout) == 24
    assert C.sizeof(_cabi.Layout) == 8 + 24 * _cabi.MAX_POPS
    assert C.sizeof(_cabi.Cond) == 8 + 8 * 8 + 8 * 8 + 4 * 8 + 8
    assert C.sizeof(_cabi.Job) == 4 * 3 + 4 * 8 + 4 + 2 * C.sizeof(_cabi.Cond) + 16
    assert C.sizeof(_cabi.HostResults) == 11 * 8


def test_layout_and_bits():
    from sai_b200 import _cabi
    from sai_b200.encode import make_layout

    lib = _cabi.load()
    assert [lib.sai_bits_for_max_value(v) for v in (1, 2, 3, 4, 6, 7, 14)] == [2, 2, 3, 3, 3, 4, 4]
    lay = make_layout([1500, 1000, 4], [2, 2, 2])
    assert [(lay.pop[i].n_groups, lay.pop[i].n_pairs, lay.pop[i].pair_off) for i in range(3)] == [(47, 47, 0), (32, 32, 47), (1, 1, 79)]
    assert lay.pairs_per_site == 80
    assert lib.sai_packed_bytes(C.byref(lay), 6_000_000) == 187_500 * 80 * 256
    assert lib.sai_num_tiles(33) == 2
    lay3 = make_layout([33, 5], [4, 3])
    assert (lay3.pop[0].bits, lay3.pop[0].n_pairs, lay3.pop[1].bits, lay3.pop[1].n_pairs) == (3, 3, 3, 2)
    with pytest.raises(ValueError, match="ploidy must be a positive integer"):
        make_layout([3], [0])
    with pytest.raises(ValueError):
        make_layout([3] * 17, [2] * 17)


@pytest.mark.parametrize("n_sites", [0, 1, 31, 32, 33, 1000])
def test_pack_unpack_roundtrip(n_sites):
    from sai_b200.encode import pack_populations, unpack_population

    rng = np.random.default_rng(n_sites)
    mats = [rng.integers(-2, 3, size=(n_sites, 70)).astype(np.int8), rng.integers(-1, 2, size=(n_sites, 1)).astype(np.int64),
            rng.integers(-1, 5, size=(n_sites, 33)).astype(np.int8), rng.integers(0, 9, size=(n_sites, 64)).astype(np.int8)]
    pg = pack_populations(mats, [2, 1, 4, 8], np.arange(n_sites))
    assert [pg.layout.pop[i].bits for i in range(4)] == [2, 2, 3, 4]
    for i, m in enumerate(mats):
        assert np.array_equal(unpack_population(pg, i), np.where(m < 0, -1, m).astype(np.int8))
    if n_sites > 40:
        assert np.array_equal(unpack_population(pg, 2, 35, 4), np.where(mats[2][35:39] < 0, -1, mats[2][35:39]))


def test_negative_table_for_dd():
    """The DD side table: raw values of every missing call, per population in
    (site, individual) order, for any integer dtype."""
    from sai_b200.encode import negative_table, pack_populations

    rng = np.random.default_rng(11)
    mats = [rng.integers(-2, 3, size=(50, n)).astype(dt) for n, dt in ((7, np.int8), (3, np.int64), (1, np.int16))]
    mats[1][4, 1] = -(2**40)  # clipped, still negative
    off, site, ind, val = negative_table(mats)
    assert off[0] == 0 and off.dtype == np.int64 and site.dtype == ind.dtype == val.dtype == np.int32
    for p, m in enumerate(mats):
        sl = slice(int(off[p]), int(off[p + 1]))
        r, c = np.nonzero(m < 0)
        assert np.array_equal(site[sl], r) and np.array_equal(ind[sl], c)
        assert np.array_equal(val[sl], np.maximum(m[r, c].astype(np.int64), -(2**31) + 1))
        key = site[sl].astype(np.int64) * 1000 + ind[sl]
        assert np.all(np.diff(key) > 0)
    pg = pack_populations(mats, [2, 2, 2], np.arange(50), keep_negatives=True)
    assert np.array_equal(pg.neg_off, off) and np.array_equal(pg.neg_val, val)
    # the native int8 scan over many site blocks and odd widths
    for width in (1, 7, 8, 37, 64):
        big = rng.integers(-3, 3, size=(6000, width)).astype(np.int8)
        big[rng.random(big.shape) < 0.7] = 0
        o, s_, i_, v_ = negative_table([big])
        r, c = np.nonzero(big < 0)
        assert o[1] == r.shape[0] and np.array_equal(s_, r) and np.array_equal(i_, c) and np.array_equal(v_, big[r, c])
    assert pack_populations(mats, [2, 2, 2], np.arange(50)).neg_off is None
    # column blocks of one matrix (what the VCF reader hands out) pack in place: same tiles, same table
    wide = rng.integers(-2, 3, size=(300, 60)).astype(np.int8)
    views = [wide[:, 0:33], wide[:, 33:40], wide[:, 40:60]]
    a = pack_populations(views, [2, 2, 2], np.arange(300), keep_negatives=True)
    b = pack_populations([np.ascontiguousarray(v) for v in views], [2, 2, 2], np.arange(300), keep_negatives=True)
    assert not views[0].flags.c_contiguous and np.array_equal(a.packed, b.packed)
    assert all(np.array_equal(getattr(a, k), getattr(b, k)) for k in ("neg_off", "neg_site", "neg_ind", "neg_val"))


def test_zt_encoder_vector_path_is_identical():
    """The AVX-512 VBMI2 record encoder (when this CPU has it) writes the bytes of the portable
    one: sparse, half-dense and dense tiles, odd pair counts, padding constants, partial last tile."""
    from sai_b200 import _cabi
    from sai_b200.encode import pack_populations

    lib = _cabi.load()
    assert lib.sai_zt_isa() in (b"avx512vbmi2", b"portable")
    rng = np.random.default_rng(77)
    for sizes, ploidy, density in (((300, 90, 4), [2, 2, 2], 0.02), ((45, 33, 7, 70), [4, 3, 1, 8], 0.2),
                                   ((64, 96, 32), [2, 2, 2], 0.9), ((1500, 1000, 4), [2, 2, 2], 0.05)):
        n = 333
        f = rng.beta(0.3, 3.0, size=n) * density * 5
        mats = [rng.binomial(p, np.clip(f, 0, 1)[:, None], size=(n, k)).astype(np.int8) for k, p in zip(sizes, ploidy)]
        mats[0][rng.random(mats[0].shape) < 0.01] = -1
        pg = pack_populations(mats, ploidy, np.arange(1, n + 1))
        outs = []
        for isa in (1, 0):
            cap = int(lib.sai_zt_bound(C.byref(pg.layout), n))
            out = np.full(cap, 0xEE, dtype=np.uint8)
            off = np.zeros(pg.n_tiles + 1, dtype=np.uint64)
            length = lib.sai_zt_encode_isa(C.byref(pg.layout), pg.packed.ctypes.data, n, out.ctypes.data, cap, off.ctypes.data, 2, isa)
            assert length >= 0
            outs.append((out[:length].copy(), off))
            assert (out[length:] == 0xEE).all()
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        back = np.empty_like(pg.packed)
        assert lib.sai_zt_decode_host(C.byref(pg.layout), outs[1][0].ctypes.data, outs[1][1].ctypes.data, n, back.ctypes.data) == 0
        assert np.array_equal(back, pg.packed)


@pytest.mark.parametrize("kind", ["sparse", "incompressible", "mixed_bits", "one_site", "empty", "views", "many_blocks"])
def test_zt_pack_i8_equals_pack_then_encode(kind):
    """`sai_zt_pack_i8` (the int8 pipeline's block encoder: pack a tile, encode it in cache, append
    the record) writes the stream and directory of `sai_zt_encode(sai_pack_i8_all(...))` byte for
    byte -- sparse and incompressible (raw-tile) data, odd plane counts, row-strided views, several
    blocks of 32 tiles, any thread count -- and reports a value that does not fit the planes."""
    from sai_b200.encode import MatrixGenotypes, compress, compress_matrices, decompress, pack_populations

    rng = np.random.default_rng(len(kind))
    ploidy = [2, 2, 2]
    if kind == "sparse":
        n, sizes = 2500, (300, 90, 4)
        f = rng.beta(0.2, 2.0, size=n)
        mats = [rng.binomial(2, f[:, None], size=(n, k)).astype(np.int8) for k in sizes]
        mats[1][rng.random(mats[1].shape) < 0.002] = -1
        mats[0][rng.random(mats[0].shape) < 0.001] = -2
    elif kind == "incompressible":
        n, sizes = 1100, (64, 96, 32)
        mats = [rng.integers(-1, 3, size=(n, k)).astype(np.int8) for k in sizes]
    elif kind == "mixed_bits":
        n, ploidy, sizes = 1300, [4, 3, 1, 8], (45, 32, 7, 70)
        f = rng.beta(0.3, 3.0, size=n)
        mats = [rng.binomial(p, f[:, None], size=(n, k)).astype(np.int8) for k, p in zip(sizes, ploidy)]
        mats[3][rng.random(mats[3].shape) < 0.01] = -2
    elif kind == "many_blocks":  # 39 blocks of raw tiles, 655 KB each: a thread's 16 MB arena chunk rolls over
        n, sizes = 40_000, (1500, 1000, 4)
        mats = [rng.integers(0, 3, size=(n, k), dtype=np.int8) for k in sizes]
    elif kind == "one_site":
        n, sizes = 1, (3, 2, 1)
        mats = [np.ones((1, k), np.int8) for k in sizes]
    elif kind == "empty":
        n, sizes = 0, (5, 3, 1)
        mats = [np.zeros((0, k), np.int8) for k in sizes]
    else:  # column blocks of one parsed matrix
        n = 2100
        f = rng.beta(0.2, 2.0, size=n)
        whole = rng.binomial(2, f[:, None], size=(n, 160)).astype(np.int8)
        whole[rng.random(whole.shape) < 0.01] = -1
        mats = [whole[:, :100], whole[:, 100:157], whole[:, 157:]]
    pos = np.arange(1, n + 1)
    pg = pack_populations(mats, ploidy, pos)
    want = compress(pg, n_threads=2)
    mg = MatrixGenotypes(pg.layout, n, pos, mats)
    for threads in (1, 3, 0):
        got = compress_matrices(mg, n_threads=threads)
        assert np.array_equal(got.tile_off, want.tile_off), (kind, threads)
        assert np.array_equal(got.stream, want.stream), (kind, threads)
    assert np.array_equal(decompress(got).packed, pg.packed)
    if kind in ("incompressible", "many_blocks"):
        assert (got.tile_off[:-1] >> np.uint64(63)).all()
    if n > 1 and kind != "many_blocks":
        from sai_b200 import _cabi

        with pytest.raises(_cabi.SaiError):
            compress_matrices(mg, out=np.empty(16, dtype=np.uint8))
        bad = [m.copy() for m in mats]
        bad[0][n // 2, 0] = 100
        if pg.layout.pop[0].bits < 7:
            with pytest.raises(ValueError, match="does not fit"):
                compress_matrices(MatrixGenotypes(pg.layout, n, pos, bad))


@pytest.mark.parametrize("shape", ["sparse", "dense", "mixed_bits", "tiny", "empty"])
def test_zt_roundtrip_host(shape):
    """encode -> host decode is the identity on the packed tiles (sparse data, dense data that
    takes the raw-tile escape, 3/4-plane populations with odd word counts, partial last tile)."""
    from sai_b200.encode import compress, decompress, pack_populations

    rng = np.random.default_rng(21)
    if shape == "sparse":
        n, sizes, ploidy = 1000, (300, 90, 4), [2, 2, 2]
        f = rng.beta(0.2, 2.0, size=n)
        mats = [rng.binomial(2, f[:, None], size=(n, k)).astype(np.int8) for k in sizes]
        mats[1][rng.random(mats[1].shape) < 0.002] = -1
    elif shape == "dense":
        n, sizes, ploidy = 333, (64, 96, 32), [2, 2, 2]
        mats = [rng.integers(-1, 3, size=(n, k)).astype(np.int8) for k in sizes]
    elif shape == "mixed_bits":
        n, ploidy = 257, [4, 3, 1, 8]
        sizes = (45, 32, 7, 70)
        f = rng.beta(0.3, 3.0, size=n)
        mats = [rng.binomial(p, f[:, None], size=(n, k)).astype(np.int8) for k, p in zip(sizes, ploidy)]
        mats[3][rng.random(mats[3].shape) < 0.01] = -2
    elif shape == "tiny":
        n, sizes, ploidy = 1, (1, 1, 1), [1, 2, 1]
        mats = [np.zeros((1, 1), np.int8) for _ in sizes]
    else:
        n, sizes, ploidy = 0, (5, 3, 1), [2, 2, 2]
        mats = [np.zeros((0, k), np.int8) for k in sizes]
    pg = pack_populations(mats, ploidy, np.arange(1, n + 1))
    zt = compress(pg, n_threads=3)
    assert zt.tile_off.shape == (pg.n_tiles + 1,) and int(zt.tile_off[-1]) == zt.stream.nbytes
    back = decompress(zt)
    assert np.array_equal(back.packed, pg.packed)
    raw = (zt.tile_off[:-1] >> np.uint64(63)).astype(bool)
    if shape == "sparse":
        assert not raw.any() and zt.stream.nbytes < 0.5 * pg.packed.nbytes
    if shape == "dense":
        assert raw.all() and zt.stream.nbytes == pg.packed.nbytes
    if shape == "tiny":  # site 0 is all hom-ref (no payload); sites 1..31 of the tile are coded all-missing
        assert not raw.any() and 0 < zt.stream.nbytes < pg.packed.nbytes
    # a buffer that is too small is reported, not overrun
    if n > 0:
        from sai_b200 import _cabi

        with pytest.raises(_cabi.SaiError):
            compress(pg, out=np.empty(16, dtype=np.uint8))


def test_pack_rejects_out_of_domain():
    from sai_b200.encode import pack_populations

    g = np.array([[0, 1, 3]], dtype=np.int8)
    with pytest.raises(ValueError, match="does not fit"):
        pack_populations([g], [2], np.arange(1), bits=[2])
    pack_populations([g], [2], np.arange(1))  # auto: 3 bit-planes


def test_job_validation_messages():
    from sai_b200.scoring import make_job

    with pytest.raises(ValueError, match=r"Parameters w must be within the range \[0, 1\]."):
        make_job(0, 1, [2], True, u=dict(w=-0.1, x=0.5, y_list=[("=", 1.0)]))
    with pytest.raises(ValueError, match="Invalid value in y_list"):
        make_job(0, 1, [2], True, q=dict(w=0.1, quantile=0.5, y_list=[("=", 1.1)]))
    with pytest.raises(ValueError, match="Invalid operator in y_list"):
        make_job(0, 1, [2], True, q=dict(w=0.1, quantile=0.5, y_list=[("==", 1.0)]))
    with pytest.raises(ValueError, match="The length of src_gts_list and y_list must match"):
        make_job(0, 1, [2, 3], True, q=dict(w=0.1, quantile=0.5, y_list=[("=", 1.0)]))
    j = make_job(0, 1, [2], False, u=dict(w=0.1, x=0.5, y_list=[("=", 0.8)]))
    assert j.u.one_minus_y[0] == 1 - 0.8 == 0.19999999999999996 and j.q.enabled == 0


def test_windows_match_oracle():
    import sai_oracle as orc
    from sai_b200 import windows as W

    rng = np.random.default_rng(1)
    for _ in range(300):
        first = int(rng.integers(0, 10**6))
        last = first + int(rng.integers(0, 10**6))
        step = int(rng.integers(1, 60000))
        size = step + int(rng.integers(0, 100000))
        start = None if rng.random() < 0.5 else int(rng.integers(0, first + 2))
        a, b = W.split_genome([first, last], size, step, start), orc.split_genome([first, last], size, step, start)
        assert a == b
        n = int(rng.integers(1, 9))
        assert W.split_windows_ranges(a, n) == orc.split_windows_ranges(b, n)
        for (s, e) in W.split_windows_ranges(a, n):
            pass
    wins = W.split_genome([2309, 48989], 10000, 5000)
    assert W.split_windows_ranges(wins, 2) == [(1, 30000), (25001, 55000)]  # reference test_chunk_generator.py:39
    # every shard re-derives exactly its own windows, in order (the halo is win_len - step)
    for n in (1, 2, 3, 4, 8):
        parts = [W.chunk_windows(s, e, 10000, 5000) for s, e in W.split_windows_ranges(wins, n)]
        assert [w for p in parts for w in p] == wins
    with pytest.raises(ValueError, match="`step_size` cannot be greater than `window_size`"):
        W.split_genome([1, 2], 20, 25)
    with pytest.raises(ValueError, match="`pos` array must not be empty"):
        W.split_genome([], 30, 10)
    with pytest.raises(ValueError, match="must be positive integers"):
        W.split_genome([1], 0, 0)


def test_config_mirror():
    from sai_b200.configs import PloidyConfig, StatConfig, parse_comparator

    assert parse_comparator("=1", "U", "src") == ("=", 1.0)
    assert parse_comparator(">=0.2", "U", "src") == (">=", 0.2)
    assert parse_comparator("<0.5", "Q", "src") == ("<", 0.5)
    with pytest.raises(ValueError, match="must contain a valid comparator"):
        parse_comparator("0.5", "U", "src")
    with pytest.raises(ValueError, match="between 0 and 1"):
        parse_comparator("=1.5", "U", "src")
    sc = StatConfig({"U": {"ref": {"A": 0.3}, "tgt": {"B": 0.5}, "src": {"C": "=1", "D": "<=0.25"}}, "fd": True})
    assert list(sc.get_parameters("U")["src"].values()) == [("=", 1.0), ("<=", 0.25)]
    with pytest.raises(ValueError, match="not supported"):
        StatConfig({"Z": True})
    with pytest.raises(ValueError, match="exactly the keys"):
        StatConfig({"U": {"ref": {"A": 0.3}}})
    pc = PloidyConfig({"ref": {"A": 2}, "tgt": {"B": 4}, "src": {"C": 1, "D": 2}})
    assert pc.get_ploidy("src") == [1, 2] and pc.get_ploidy("tgt", "B") == 4 and pc.get_ploidy("outgroup") is None
    with pytest.raises(ValueError, match="positive integer"):
        PloidyConfig({"ref": {"A": 0}, "tgt": {"B": 4}, "src": {"C": 1}})
    with pytest.raises(KeyError):
        pc.get_ploidy("ref", "nope")


def test_vcf_ingest_semantics(tmp_path):
    """numbers={'GT': ploidy} padding / truncation, '.' alleles, region filter,
    ancestral-allele filter and flip (sai/utils/utils.py:78-186, 492-555)."""
    from sai_b200.configs import PloidyConfig
    from sai_b200.vcf import read_data

    vcf = tmp_path / "t.vcf"
    vcf.write_text(
        "##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\tc\n"
        "1\t10\t.\tA\tT\t.\t.\t.\tGT\t0|1\t1/1\t.|.\n"
        "1\t20\t.\tA\tT\t.\t.\t.\tGT:DP\t1|.:3\t0|0|1|1:4\t1:2\n"
        "1\t30\t.\tG\tC\t.\t.\t.\tGT\t0|0\t0|1\t1|1\n"
        "1\t40\t.\tG\tC\t.\t.\t.\tGT\t0|0\t0|1\t1|1\n"
        "2\t10\t.\tA\tT\t.\t.\t.\tGT\t1|1\t1|1\t1|1\n"
    )
    for g, pops in (("ref", "R a"), ("tgt", "T b"), ("src", "S c")):
        (tmp_path / f"{g}.list").write_text(pops.replace(" ", "\t") + "\n")
    anc = tmp_path / "anc.bed"
    anc.write_text("1\t9\t10\tA\n1\t19\t20\tT\n1\t29\t30\tN\n2\t9\t10\tA\n")  # pos 30: neither REF nor ALT; pos 40: absent
    lists = [str(tmp_path / f"{g}.list") for g in ("ref", "tgt", "src")]
    pc = PloidyConfig({"ref": {"R": 2}, "tgt": {"T": 4}, "src": {"S": 2}})
    d = read_data(str(vcf), "1", pc, *lists, None, None)
    assert d["ref"][0]["R"].POS.tolist() == [10, 20, 30, 40] and d["outgroup"] == (None, None)
    assert d["ref"][0]["R"].GT[:, 0].tolist() == [1, 0, 0, 0]  # "1|." -> 1 + (-1) = 0: called (SURVEY hard part 4)
    assert d["tgt"][0]["T"].GT[:, 0].tolist() == [0, 2, -1, -1]  # diploid read as ploidy 4: padded with -1
    assert d["src"][0]["S"].GT[:, 0].tolist() == [-2, 0, 2, 2]  # ".|." -> -2; haploid "1" -> 1 + (-1)
    d = read_data(str(vcf), "1", pc, *lists, None, str(anc), start=5, end=35)
    assert d["ref"][0]["R"].POS.tolist() == [10, 20]  # 30 removed (anc not REF/ALT), 40 outside region
    assert d["ref"][0]["R"].GT[:, 0].tolist() == [1, 2]  # pos 20 flipped: alleles (1,-1) -> (0,2)
    assert d["src"][0]["S"].GT[:, 0].tolist() == [-2, 2]  # flipped haploid "1": (1,-1) -> (0, 2)
    assert read_data(str(vcf), "3", pc, *lists, None, None)["ref"][0] is None


def _random_vcf(path, rng, n_sites, n_samples, crlf=False, final_newline=True):
    lines = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n_samples))]
    pos = np.sort(rng.choice(np.arange(1, 50 * n_sites), size=n_sites, replace=False))
    bases = "ACGT"
    recs = []
    for p in pos:
        chrom = "7" if rng.random() < 0.85 else "77"  # "77" must not match chromosome "7"
        ref = bases[rng.integers(4)]
        alt = bases[(bases.index(ref) + 1 + rng.integers(3)) % 4] + (",<DEL>" if rng.random() < 0.1 else "")
        fmt = rng.choice(["GT", "GT:DP", "DP:GT:GQ"])
        fields = []
        for _ in range(n_samples):
            pl = int(rng.choice([1, 2, 2, 2, 4]))
            alleles = [("." if rng.random() < 0.08 else str(int(rng.integers(0, 3 if rng.random() < 0.05 else 2)))) for _ in range(pl)]
            gt = alleles[0]
            for a in alleles[1:]:
                gt += ("|" if rng.random() < 0.5 else "/") + a
            fields.append({"GT": gt, "GT:DP": gt + ":12", "DP:GT:GQ": "7:" + gt + ":30"}[fmt])
        lines.append("\t".join([chrom, str(p), ".", ref, alt, ".", "PASS", "AA=x", fmt] + fields))
        recs.append((chrom, int(p), ref, alt.split(",")[0]))
    text = ("\r\n" if crlf else "\n").join(lines) + (("\r\n" if crlf else "\n") if final_newline else "")
    open(path, "w", newline="").write(text)
    return recs


@pytest.mark.parametrize("crlf, final_newline", [(False, True), (True, True), (False, False)])
def test_native_vcf_reader_matches_python_reader(tmp_path, crlf, final_newline):
    """sai_vcf_parse_gt (one native pass for all populations) == the pure-Python
    reader, with and without an ancestral-allele table and a region."""
    from sai_b200.configs import PloidyConfig
    from sai_b200.vcf import read_data

    rng = np.random.default_rng(17)
    vcf = tmp_path / "r.vcf"
    recs = _random_vcf(vcf, rng, 400, 23, crlf, final_newline)
    (tmp_path / "ref.list").write_text("".join(f"R1\ts{i}\n" for i in range(0, 9)) + "".join(f"R2\ts{i}\n" for i in range(5, 12)))
    (tmp_path / "tgt.list").write_text("".join(f"T\ts{i}\n" for i in range(12, 20)))
    (tmp_path / "src.list").write_text("S1\ts20\nS1\ts21\nS2\ts22\n")
    anc = tmp_path / "anc.bed"
    with open(anc, "w") as f:
        for chrom, p, ref, alt in recs:
            r = rng.random()
            if r < 0.1:
                continue  # no ancestral allele: record dropped
            allele = ref if r < 0.5 else (alt if r < 0.9 else "N")
            f.write(f"{chrom}\t{p - 1}\t{p}\t{allele}\n")
    pc = PloidyConfig({"ref": {"R1": 2, "R2": 4}, "tgt": {"T": 2}, "src": {"S1": 1, "S2": 3}})
    lists = [str(tmp_path / f"{g}.list") for g in ("ref", "tgt", "src")]
    for anc_file in (None, str(anc)):
        for region in ((None, None), (recs[40][1], recs[300][1])):
            a = read_data(str(vcf), "7", pc, *lists, None, anc_file, start=region[0], end=region[1], native=True)
            b = read_data(str(vcf), "7", pc, *lists, None, anc_file, start=region[0], end=region[1], native=False)
            n_rows = None
            for g in ("ref", "tgt", "src"):
                assert list(a[g][0]) == list(b[g][0]) and a[g][1] == b[g][1]
                for p in a[g][0]:
                    assert np.array_equal(a[g][0][p].POS, b[g][0][p].POS), (g, p)
                    assert np.array_equal(a[g][0][p].GT, b[g][0][p].GT), (g, p)
                    n_rows = a[g][0][p].POS.size
            assert n_rows and n_rows > 20
    with pytest.raises(ValueError, match="Failed to read VCF"):
        (tmp_path / "bad.list").write_text("R1\tnobody\nR2\ts1\n")
        read_data(str(vcf), "7", pc, str(tmp_path / "bad.list"), lists[1], lists[2], None, None)


def test_native_vcf_parser_segments_and_row_cap():
    """Pass 1 of the native parser cuts the text into per-thread segments at line starts and the
    caller may offer fewer output rows than there are records: resuming at bytes_consumed gives
    the same rows as one big call, whatever the thread count."""
    import ctypes as C

    from sai_b200 import _cabi

    lib = _cabi.load()
    rng = np.random.default_rng(8)
    n_rec, n_smp = 420, 2000
    tok = np.array(["0|0", "0|1", "1|0", "1|1", ".|.", "0/1", "1|.", "10|1"])
    pos = np.cumsum(rng.integers(1, 50, size=n_rec))
    lines = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n_smp))]
    for i in range(n_rec):
        chrom = "2" if i % 17 == 5 else "1"  # other chromosomes are skipped
        g = rng.choice(len(tok), size=n_smp, p=[.6, .1, .1, .1, .03, .03, .02, .02])
        lines.append(f"{chrom}\t{pos[i]}\t.\tA\tG\t.\t.\t.\tGT\t" + "\t".join(tok[g]))
    text = ("\n".join(lines) + "\n").encode() + b"1\t999999\t.\tA\tG\t.\t.\t.\tGT\t0|0"  # incomplete last line
    assert len(text) > 3 << 20
    cols = np.ascontiguousarray(rng.permutation(n_smp)[:300], dtype=np.int32)
    pl = np.ascontiguousarray(rng.choice([1, 2, 3], size=300), dtype=np.int32)

    def parse(cap, n_threads):
        at, pos_parts, gt_parts = 0, [], []
        consumed = C.c_int64(0)
        while True:
            out_pos = np.empty(cap, dtype=np.int32)
            out_gt = np.empty((cap, 300), dtype=np.int8)
            n = lib.sai_vcf_parse_gt(C.c_char_p(text[at:]), len(text) - at, b"1", 1, 0, cols.ctypes.data, pl.ctypes.data, 300,
                                     None, None, 0, out_pos.ctypes.data, out_gt.ctypes.data, 300, cap, C.byref(consumed), n_threads)
            assert n >= 0
            pos_parts.append(out_pos[:n].copy())
            gt_parts.append(out_gt[:n].copy())
            if consumed.value == 0:
                break
            at += consumed.value
        return np.concatenate(pos_parts), np.concatenate(gt_parts), at

    p1, g1, at1 = parse(10_000, 1)
    assert p1.shape[0] == sum(1 for i in range(n_rec) if i % 17 != 5)
    assert text[at1:].startswith(b"1\t999999")  # the incomplete line is left for the next call
    for cap, nt in ((10_000, 4), (37, 3), (1, 8)):
        p2, g2, at2 = parse(cap, nt)
        assert np.array_equal(p1, p2) and np.array_equal(g1, g2) and at1 == at2, (cap, nt)
    # spot-check values against the token table
    assert set(np.unique(g1)) <= set(range(-3, 23))


def test_bgzf_parallel_inflate_and_reader(tmp_path):
    """bgzip input: blocks are indexed and inflated in parallel by the native library; reading a
    bgzipped VCF gives exactly what the plain-text and the plain-gzip files give, for batch sizes
    that cut lines in the middle; corrupt blocks are reported."""
    import ctypes as C
    import gzip

    from sai_b200 import _cabi
    from sai_b200.vcf import _is_bgzf, _native_read, write_bgzf

    lib = _cabi.load()
    rng = np.random.default_rng(5)
    n_rec, n_smp = 600, 300
    tok = np.array(["0|0", "0|1", "1|0", "1|1", ".|.", "0/1"])
    pos = np.cumsum(rng.integers(1, 50, size=n_rec))
    lines = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n_smp))]
    for i in range(n_rec):
        g = rng.choice(len(tok), size=n_smp, p=[.6, .1, .1, .1, .05, .05])
        lines.append(f"1\t{pos[i]}\t.\tA\tG\t.\t.\t.\tGT\t" + "\t".join(tok[g]))
    text = ("\n".join(lines) + "\n").encode()
    plain, bgz, gz = tmp_path / "a.vcf", tmp_path / "a.vcf.gz", tmp_path / "b.vcf.gz"
    plain.write_bytes(text)
    write_bgzf(str(bgz), text, block=4000)
    with gzip.open(gz, "wb") as f:
        f.write(text)
    assert _is_bgzf(str(bgz)) and not _is_bgzf(str(gz)) and not _is_bgzf(str(plain))
    # raw API: scan + inflate == the text
    raw = np.frombuffer(bgz.read_bytes(), dtype=np.uint8)
    nb = 1 << 12
    block_off, out_off = np.empty(nb, np.int64), np.empty(nb + 1, np.int64)
    consumed = C.c_int64(0)
    n = lib.sai_bgzf_scan(raw.ctypes.data, raw.size, nb, 1 << 40, block_off.ctypes.data, out_off.ctypes.data, C.byref(consumed))
    assert n == -(-len(text) // 4000) + 1 and consumed.value == raw.size and out_off[n] == len(text)
    out = np.empty(len(text), np.uint8)
    _cabi.check(lib.sai_bgzf_inflate(raw.ctypes.data, block_off.ctypes.data, out_off.ctypes.data, n, out.ctypes.data, 3))
    assert out.tobytes() == text
    bad = raw.copy()
    bad[int(block_off[5]) + 30] ^= 0xFF
    with pytest.raises(ValueError, match="corrupt"):
        _cabi.check(lib.sai_bgzf_inflate(bad.ctypes.data, block_off.ctypes.data, out_off.ctypes.data, n, out.ctypes.data, 2))
    # the reader: same rows from the three containers, for several batch sizes
    req = [(f"s{i}", 2) for i in rng.permutation(n_smp)[:40]]
    p0, g0 = _native_read(str(plain), "1", None, None, req, None)
    assert p0.shape[0] == n_rec
    for path, kw in ((gz, {}), (bgz, {}), (bgz, dict(group_blocks=1)), (bgz, dict(fused_bgzf=False)),
                     (bgz, dict(fused_bgzf=False, batch_bytes=10_000)), (bgz, dict(fused_bgzf=False, batch_bytes=1))):
        p1, g1 = _native_read(str(path), "1", None, None, req, None, **kw)
        assert np.array_equal(p0, p1) and np.array_equal(g0, g1), (path, kw)
    # no trailing newline + region filter
    write_bgzf(str(bgz), text[:-1], block=3000)
    p1, g1 = _native_read(str(bgz), "1", int(pos[10]), int(pos[500]), req, None, batch_bytes=50_000)
    keep = (p0 >= pos[10]) & (p0 <= pos[500])
    assert np.array_equal(p0[keep], p1) and np.array_equal(g0[keep], g1)


def test_chromosome_span_native(tmp_path):
    """First / last POS and record count of a chromosome (ChunkGenerator.__init__,
    chunk_generator.py:64-76) from one native scan, for the three containers, a chromosome
    name that is a prefix of another one, and a file without a trailing newline."""
    import gzip

    from sai_b200.vcf import chromosome_span, write_bgzf

    rng = np.random.default_rng(9)
    head = "##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts0\ts1\n"
    recs, expect = [], {}
    for chrom, n in (("1", 4000), ("11", 2500), ("X", 1)):
        pos = np.cumsum(rng.integers(1, 90, size=n))
        expect[chrom] = (int(pos[0]), int(pos[-1]), n)
        recs += [f"{chrom}\t{p}\t.\tA\tG\t.\t.\t.\tGT\t0|1\t" + "1|1" * 300 for p in pos]  # long lines: > 1 MB segments
    text = (head + "\n".join(recs)).encode()  # no trailing newline
    assert len(text) > 4 << 20
    plain, bgz, gz = tmp_path / "a.vcf", tmp_path / "a.bgz.vcf.gz", tmp_path / "a.vcf.gz"
    plain.write_bytes(text)
    write_bgzf(str(bgz), text)
    with gzip.open(gz, "wb") as f:
        f.write(text)
    for path in (plain, bgz, gz):
        for chrom, want in expect.items():
            assert chromosome_span(str(path), chrom, n_threads=3) == want, (path, chrom)
        assert chromosome_span(str(path), "2") is None
    # one-chromosome files are answered from their two ends (record count -1 = not counted)
    only = [r for r in recs if r.startswith("11\t")]
    for trailing in (b"\n", b""):
        text1 = (head + "\n".join(only)).encode() + trailing
        plain.write_bytes(text1)
        write_bgzf(str(bgz), text1, block=5000)
        for path in (plain, bgz):
            assert chromosome_span(str(path), "11") == expect["11"][:2] + (-1,), (path, trailing)
            assert chromosome_span(str(path), "1") is None


def test_chunk_generator_mirror(tmp_path):
    """ChunkGenerator(vcf, chr, step, window, num_chunks): the reference's constructor order,
    ``get()`` dicts and the window-range chunks with their overlap
    (tests/generators/test_chunk_generator.py:25-40: chr21 2309..48989, 10 kb / 5 kb, 2 chunks
    -> [(1, 30000), (25001, 55000)])."""
    from sai_b200.generators import ChunkGenerator

    vcf_path = tmp_path / "t.vcf"
    rows = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts0"]
    rows += [f"20\t{p}\t.\tA\tG\t.\t.\t.\tGT\t0|1" for p in (5, 77)]
    rows += [f"21\t{p}\t.\tA\tG\t.\t.\t.\tGT\t0|1" for p in (2309, 9000, 31000, 48989)]
    vcf_path.write_text("\n".join(rows) + "\n")
    gen = ChunkGenerator(vcf_file=str(vcf_path), chr_name="21", window_size=10000, step_size=5000, num_chunks=2)
    assert len(gen) == 2
    assert list(gen.get()) == [{"chr_name": "21", "start": 1, "end": 30000}, {"chr_name": "21", "start": 25001, "end": 55000}]
    # more chunks than windows: empty ranges are dropped (chunk_generator.py:136-139)
    assert len(ChunkGenerator(str(vcf_path), "20", 1000, 1000, 5)) == 1
    with pytest.raises(ValueError, match="Chromosome 22 not found in VCF."):
        ChunkGenerator(str(vcf_path), "22", 5000, 10000, 2)


def test_mp_pool_contract():
    """mp_pool (sai/multiprocessing/mp_pool.py:43-73; reference test tests/multiprocessing/test_mp_pool.py:25-46):
    every parameter dict goes through ``run`` in a spawned worker, results reach ``process_items`` in generator order,
    every worker is pinned to one of the given devices, and an exception in ``run`` surfaces in the caller."""
    from helpers import ListGenerator, SquareProcessor

    from sai_b200.multiprocessing import mp_pool, mp_worker

    assert mp_worker((SquareProcessor(), {"x": 3})) == [(9, None)]
    proc = SquareProcessor()
    mp_pool(proc, ListGenerator([1, 2, 3, 4, 5]), nprocess=2, devices=[0, 1])
    assert [v for v, _ in proc.final_results] == [1, 4, 9, 16, 25]
    assert {d for _, d in proc.final_results} <= {0, 1}
    with pytest.raises(ValueError, match="cannot process -2"):
        mp_pool(SquareProcessor(), ListGenerator([1, -2, 3]), nprocess=2, devices=[0])
    with pytest.raises(RuntimeError, match="needs a CUDA device"):
        mp_pool(SquareProcessor(), ListGenerator([1]), nprocess=1, devices=[])


@pytest.mark.parametrize("n_smp, block", [(120, 3000), (700, 1000)])  # lines shorter / longer than a bgzip block
def test_sorted_region_read_without_index(tmp_path, n_smp, block):
    """Region reads of a one-chromosome sorted file (plain and bgzip) touch only the byte range of the region --
    found by bisection, no index file -- and return exactly the rows of a full scan filtered by POS, for regions at
    the start, in the middle, at the end, empty ones and single positions; other files fall back to the full scan."""
    import gzip

    from sai_b200 import vcf as V

    rng = np.random.default_rng(12)
    n_rec = 3000 if n_smp < 200 else 600
    tok = np.array(["0|0", "0|1", "1|0", "1|1", ".|."])
    pos = np.cumsum(rng.integers(1, 40, size=n_rec))
    head = "##fileformat=VCFv4.1\n##contig=<ID=7>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n_smp)) + "\n"
    recs = [f"7\t{p}\t.\tA\tG\t.\t.\t.\tGT\t" + "\t".join(tok[rng.choice(5, size=n_smp, p=[.7, .1, .1, .05, .05])]) for p in pos]
    text = (head + "\n".join(recs)).encode()  # no trailing newline
    plain, bgz = tmp_path / "r.vcf", tmp_path / "r.vcf.gz"
    plain.write_bytes(text)
    V.write_bgzf(str(bgz), text, block=block)  # hundreds of blocks: the bisection has something to do
    req = [(f"s{i}", 2) for i in (5, 0, 77, 119)]
    p_all, g_all = V._native_read(str(plain), "7", None, None, req, None)
    assert p_all.shape[0] == n_rec
    calls = []
    orig = V._scan_text
    V._scan_text = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        regions = [(1, int(pos[10])), (int(pos[0]), int(pos[0])), (int(pos[n_rec // 2]), int(pos[n_rec // 2 + 200])),
                   (int(pos[n_rec - 10]), 10**9), (int(pos[-1]), int(pos[-1])), (int(pos[100]) + 1, int(pos[101]) - 1),
                   (10**9, 2 * 10**9), (1, 10**9), (int(pos[77]), int(pos[77]))]
        for path in (plain, bgz):
            for a, b in regions:
                p1, g1 = V._native_read(str(path), "7", a, b, req, None, batch_bytes=200_000)
                keep = (p_all >= a) & (p_all <= b)
                assert np.array_equal(p1, p_all[keep]) and np.array_equal(g1, g_all[keep]), (path, a, b)
        assert not calls  # every region was served by the bisection path
        # a second chromosome in the file: not eligible, full scan (still correct)
        two = tmp_path / "two.vcf"
        two.write_bytes(text + b"\n8\t5\t.\tA\tG\t.\t.\t.\tGT\t" + b"\t".join([b"0|1"] * n_smp) + b"\n")
        p1, g1 = V._native_read(str(two), "7", int(pos[5]), int(pos[50]), req, None)
        assert calls and np.array_equal(p1, p_all[5:51])
    finally:
        V._scan_text = orig


# ---------------------------------------------------------------- N2 pinned to the reference's own ingest KATs
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _testdata_paths():
    return (os.path.join(GOLDEN, "vcf_testdata.vcf"), os.path.join(GOLDEN, "vcf_testdata.anc.bed"),
            *[os.path.join(GOLDEN, f"vcf_testdata.{g}.list") for g in ("ref", "tgt", "src")])


def test_reference_kat_read_anc_allele():
    """tests/utils/test_utils.py:204-209: `read_anc_allele(test.anc.allele.bed, "21")` ==
    {"21": {2309: "G", 7879: "A", 11484: "-", 48989: "C"}} -- the "-" entry is kept as is."""
    from sai_b200.vcf import read_anc_allele, read_anc_arrays

    _, anc, *_ = _testdata_paths()
    assert read_anc_allele(anc, "21") == {2309: "G", 7879: "A", 11484: "-", 48989: "C"}
    arr = read_anc_arrays(anc, "21", 5000, 20000)
    assert arr.as_dict() == {7879: "A", 11484: "-"}
    with pytest.raises(ValueError, match="No ancestral allele is found for chromosome 22."):
        read_anc_allele(anc, "22")
    with pytest.raises(ValueError, match="in the region 3000-4000"):
        read_anc_allele(anc, "21", 3000, 4000)


@pytest.mark.parametrize("native", [True, False])
def test_reference_kat_check_anc_allele(native):
    """tests/utils/test_utils.py:269-318 (`test_check_anc_allele`): test.data.vcf filtered and
    polarised by test.anc.allele.bed keeps positions [2309, 7879, 48989] (11484 has "-": neither
    REF nor ALT -> removed) and flips 7879 (ancestral == ALT "A").  The reference's test holds the
    haplotype matrices (is_phased=True default: `.reshape(3, 4)`); the scoring path uses
    is_phased=False (window_generator.py:113), i.e. the per-individual sums of the same literals:
        ref1 [[0,0],[0,0]] [[1,1],[1,1]] [[0,0],[0,0]]  -> [[0,0],[2,2],[0,0]]
        tgt1 [[1,0],[0,0]] [[1,1],[1,0]] [[0,0],[0,1]]  -> [[1,0],[2,1],[0,1]]
        tgt2 [[0,0],[0,0]] [[1,1],[1,1]] [[0,0],[0,0]]  -> [[0,0],[2,2],[0,0]]"""
    from sai_b200.configs import PloidyConfig
    from sai_b200.vcf import read_data

    vcf, anc, ref_l, tgt_l, src_l = _testdata_paths()
    pc = PloidyConfig({"ref": {"ref1": 2}, "tgt": {"tgt1": 2, "tgt2": 2}, "src": {"src1": 2, "src2": 2}})
    d = read_data(vcf, "21", pc, ref_l, tgt_l, None, None, anc, native=native)
    exp_pos = [2309, 7879, 48989]
    assert np.array_equal(d["ref"][0]["ref1"].GT, [[0, 0], [2, 2], [0, 0]])
    assert np.array_equal(d["tgt"][0]["tgt1"].GT, [[1, 0], [2, 1], [0, 1]])
    assert np.array_equal(d["tgt"][0]["tgt2"].GT, [[0, 0], [2, 2], [0, 0]])
    for g, p in (("ref", "ref1"), ("tgt", "tgt1"), ("tgt", "tgt2")):
        assert np.array_equal(d[g][0][p].POS, exp_pos)
    assert d["src"] == (None, None)
    # tests/utils/test_utils.py:154-200 (`test_read_data_from_file`): without the table all 19
    # records of chromosome 21 come back (`.reshape(19, 4)`), the chromosome-20 line does not;
    # the sample dicts are those of parse_ind_file
    d = read_data(vcf, "21", pc, ref_l, tgt_l, None, None, None, native=native)
    assert d["ref"][0]["ref1"].GT.shape == (19, 2) and d["tgt"][0]["tgt2"].GT.shape == (19, 2)
    assert d["ref"][1] == {"ref1": ["ind5", "ind6"]} and d["tgt"][1] == {"tgt1": ["ind1", "ind2"], "tgt2": ["ind3", "ind4"]}
    assert d["tgt"][0]["tgt1"].POS[0] == 2309 and d["tgt"][0]["tgt1"].GT[0].tolist() == [1, 0]


def test_reference_kat_flip_snps():
    """tests/utils/test_utils.py:404-419 (`test_flip_snps`): flipping positions 100, 200, 400 of
        100 [[0,1],[1,1]]  200 [[1,1],[1,1]]  300 [[0,0],[0,1]]  400 [[1,0],[0,0]]
    gives [[1,0],[0,0]], [[0,0],[0,0]], [[0,0],[0,1]] (unchanged), [[0,1],[1,1]] -- here through
    the native parser's flip (ancestral allele == ALT) and the pure-Python `_polarise`, as sums."""
    from sai_b200.configs import PloidyConfig
    from sai_b200.vcf import read_data

    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        vcf = os.path.join(tmp, "f.vcf")
        rows = {100: ("0|1", "1|1"), 200: ("1|1", "1|1"), 300: ("0|0", "0|1"), 400: ("1|0", "0|0")}
        with open(vcf, "w") as f:
            f.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\n")
            for p, (a, b) in rows.items():
                f.write(f"1\t{p}\t.\tA\tC\t.\t.\t.\tGT\t{a}\t{b}\n")
        bed = os.path.join(tmp, "anc.bed")
        with open(bed, "w") as f:
            for p in rows:
                f.write(f"1\t{p - 1}\t{p}\t{'A' if p == 300 else 'C'}\n")  # ALT ancestral at 100, 200, 400
        lst = os.path.join(tmp, "p.list")
        open(lst, "w").write("P\ta\nP\tb\n")
        pc = PloidyConfig({"ref": {"P": 2}, "tgt": {"P": 2}, "src": {"P": 2}})
        for native in (True, False):
            d = read_data(vcf, "1", pc, lst, lst, lst, None, bed, native=native)
            assert d["ref"][0]["P"].GT.tolist() == [[1, 0], [0, 0], [0, 1], [1, 2]], native
            assert d["ref"][0]["P"].POS.tolist() == [100, 200, 300, 400]


# ---------------------------------------------------------------- advisor findings (round 1)
def _gap_vcf(tmp_path, contig="9"):
    """Sites at 1000..4000 and 60000..64000: a gap in between (a centromere)."""
    vcf = tmp_path / "gap.vcf"
    pos = list(range(1000, 4001, 10)) + list(range(60000, 64001, 10))
    with open(vcf, "w") as f:
        f.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\tc\n")
        for i, p in enumerate(pos):
            f.write(f"{contig}\t{p}\t.\tA\tC\t.\t.\t.\tGT\t{i % 2}|0\t0|{(i // 2) % 2}\t1|1\n")
    bed = tmp_path / "gap.bed"
    with open(bed, "w") as f:
        for p in pos:
            f.write(f"{contig}\t{p - 1}\t{p}\tA\n")
    for g, s in (("ref", "a"), ("tgt", "b"), ("src", "c"), ("out", "a")):
        (tmp_path / f"{g}.list").write_text(f"{g.upper()}\t{s}\n")
    return str(vcf), str(bed), pos


@pytest.mark.parametrize("native", [True, False])
def test_gap_chunk_with_anc_alleles_is_empty_not_an_error(tmp_path, native):
    """A chunk without records AND without ancestral alleles: the reference returns None from the
    VCF read before it opens the table (utils.py:123-141) -> NaN rows; only a region WITH records
    and no ancestral allele raises."""
    from sai_b200.configs import PloidyConfig
    from sai_b200.vcf import read_data

    vcf, bed, pos = _gap_vcf(tmp_path)
    pc = PloidyConfig({"ref": {"REF": 2}, "tgt": {"TGT": 2}, "src": {"SRC": 2}})
    lists = [str(tmp_path / f"{g}.list") for g in ("ref", "tgt", "src")]
    d = read_data(vcf, "9", pc, *lists, None, bed, start=20001, end=50000, native=native)
    for g in ("ref", "tgt", "src"):
        assert d[g][0] is None and d[g][1] is not None
    d = read_data(vcf, "9", pc, *lists, None, bed, start=1, end=50000, native=native)
    assert d["ref"][0]["REF"].POS.size == 301
    short = tmp_path / "short.bed"  # ancestral alleles only for the second block
    short.write_text("".join(f"9\t{p - 1}\t{p}\tA\n" for p in pos if p >= 60000))
    with pytest.raises(ValueError, match="No ancestral allele is found for chromosome 9 in the region 1-50000"):
        read_data(vcf, "9", pc, *lists, None, str(short), start=1, end=50000, native=native)
    d = read_data(vcf, "9", pc, *lists, None, str(short), start=20001, end=50000, native=native)
    assert d["tgt"][0] is None


def test_sorted_region_read_with_long_contig_name(tmp_path):
    """The text bisection reads POS up to the second TAB, whatever the length of CHROM."""
    from sai_b200.configs import PloidyConfig
    from sai_b200.vcf import read_data

    contig = "scaffold_" + "x" * 60  # 69 characters
    vcf, bed, pos = _gap_vcf(tmp_path, contig)
    pc = PloidyConfig({"ref": {"REF": 2}, "tgt": {"TGT": 2}, "src": {"SRC": 2}})
    lists = [str(tmp_path / f"{g}.list") for g in ("ref", "tgt", "src")]
    for region in ((2000, 3000), (3500, 61000), (1, 100000), (4001, 59999)):
        a = read_data(vcf, contig, pc, *lists, None, None, start=region[0], end=region[1], native=True)
        b = read_data(vcf, contig, pc, *lists, None, None, start=region[0], end=region[1], native=False)
        want = [p for p in pos if region[0] <= p <= region[1]]
        if not want:
            assert a["ref"][0] is None and b["ref"][0] is None
            continue
        assert a["ref"][0]["REF"].POS.tolist() == want == b["ref"][0]["REF"].POS.tolist()
        assert np.array_equal(a["tgt"][0]["TGT"].GT, b["tgt"][0]["TGT"].GT)


def test_empty_items_carry_the_outgroup(tmp_path):
    """window_generator.py:269-273: the rows of an empty region are the product over the outgroup
    populations too and name them in the Outgroup column."""
    from sai_b200.configs import PloidyConfig, StatConfig
    from sai_b200.preprocessors import ChunkPreprocessor

    vcf, bed, _ = _gap_vcf(tmp_path)
    pc = PloidyConfig({"ref": {"REF": 2}, "tgt": {"TGT": 2}, "src": {"SRC": 2}, "outgroup": {"OUT": 2}})
    sc = StatConfig({"U": {"ref": {"REF": 0.5}, "tgt": {"TGT": 0.1}, "src": {"SRC": "=1"}}, "fd": True})
    pre = ChunkPreprocessor(vcf, *[str(tmp_path / f"{g}.list") for g in ("ref", "tgt", "src", "out")], 10000, 10000,
                            str(tmp_path / "o.tsv"), pc, sc, anc_allele_file=bed)
    items = pre.run("9", 20001, 50000)  # no record in the region: no GPU work
    assert len(items) == 3 and [it["out_pop"] for it in items] == ["OUT"] * 3
    assert all(np.isnan(it["U"]) and np.isnan(it["fd"]) and it["nsnps"] == 0 for it in items)


@pytest.mark.parametrize("bits, vmax", [(2, 2), (3, 6), (4, 14), (5, 30), (7, 126), (8, 127)])
def test_packer_vector_paths_are_identical(bits, vmax):
    """Portable, SSE2, AVX2, AVX-512BW and AVX-512 GFNI row packers (whichever this CPU has) write the same bytes,
    for full and partial 32-/64-individual groups, row-strided views and padding sites, and all
    report an out-of-domain value."""
    from sai_b200 import _cabi
    from sai_b200.encode import make_layout

    lib = _cabi.load()
    assert lib.sai_pack_isa() in (b"avx512gfni", b"avx512bw", b"avx2", b"sse2", b"portable")
    rng = np.random.default_rng(bits)
    n_ind = [1, 31, 32, 33, 63, 64, 65, 100, 257]
    n_sites = 70
    whole = rng.integers(0, vmax + 1, size=(n_sites, sum(n_ind) + 5)).astype(np.int8)
    whole[rng.random(whole.shape) < 0.1] = -1
    whole[rng.random(whole.shape) < 0.05] = -2
    lay = make_layout(n_ind, [1] * len(n_ind), [bits] * len(n_ind))
    nbytes = int(lib.sai_packed_bytes(C.byref(lay), n_sites))
    outs = []
    for isa in (1, 2, 3, 4, 5, 0):
        out = np.full(nbytes, 0xAB, dtype=np.uint8)
        at = 0
        for p, n in enumerate(n_ind):
            view = whole[:, at : at + n]  # row stride = whole row
            assert lib.sai_pack_i8_isa(C.byref(lay), p, view.ctypes.data, n_sites, whole.strides[0], out.ctypes.data, 3, isa) == 0
            at += n
        outs.append(out)
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
    if bits == 8:
        return  # every non-negative int8 fits 8 planes
    bad = whole.copy()
    bad[17, 70] = vmax + 1  # population 3 holds columns 64..96
    for isa in (1, 2, 3, 4, 5):
        rc = lib.sai_pack_i8_isa(C.byref(lay), 3, bad[:, 64:].ctypes.data, n_sites, bad.strides[0], outs[0].ctypes.data, 1, isa)
        assert rc == _cabi.E_DOMAIN, isa


@pytest.mark.parametrize("n_ind, bits", [([1500, 1000, 4], 2), ([257, 100, 65, 33], 3), ([700, 90], 4), ([40], 2), ([300, 300], 8)])
def test_all_populations_packer_matches_per_population_packer(n_ind, bits):
    """`sai_pack_i8_all` (the int8 pipeline's packer: site-by-site over all populations, 8 sites per
    output cache line, vector line writer, non-temporal stores when the output is 64-byte aligned)
    writes exactly the bytes of the portable per-population packer -- aligned and unaligned output,
    site counts that end inside a tile and inside an 8-site batch, 1..n threads."""
    from sai_b200 import _cabi
    from sai_b200.encode import make_layout

    lib = _cabi.load()
    rng = np.random.default_rng(sum(n_ind) + bits)
    lay = make_layout(n_ind, [1] * len(n_ind), [bits] * len(n_ind))
    vmax = min((1 << bits) - 2, 127)
    for n_sites in (1, 31, 32, 37, 200, 1029):
        whole = rng.integers(0, vmax + 1, size=(n_sites, sum(n_ind) + 3)).astype(np.int8)
        whole[rng.random(whole.shape) < 0.1] = -1
        nbytes = int(lib.sai_packed_bytes(C.byref(lay), n_sites))
        want = np.full(nbytes, 0xAB, dtype=np.uint8)
        cols = np.cumsum([0] + n_ind)
        for p in range(len(n_ind)):
            assert lib.sai_pack_i8_isa(C.byref(lay), p, whole.ctypes.data + int(cols[p]), n_sites, whole.strides[0], want.ctypes.data, 1, 1) == 0
        ptrs = (C.c_void_p * len(n_ind))(*[whole.ctypes.data + int(cols[p]) for p in range(len(n_ind))])
        strides = (C.c_int64 * len(n_ind))(*[whole.strides[0]] * len(n_ind))
        raw = np.full(nbytes + 128, 0xCD, dtype=np.uint8)
        base = (-raw.ctypes.data) % 64
        for shift, threads in ((0, 1), (0, 5), (8, 3), (3, 2)):
            raw[:] = 0xCD
            got = raw[base + shift : base + shift + nbytes]
            assert lib.sai_pack_i8_all(C.byref(lay), ptrs, strides, n_sites, got.ctypes.data, threads) == 0
            assert np.array_equal(got, want), (n_sites, shift, threads)
            assert (raw[: base + shift] == 0xCD).all() and (raw[base + shift + nbytes :] == 0xCD).all()
    if bits < 8:
        whole[n_sites // 2, 1] = vmax + 1
        assert lib.sai_pack_i8_all(C.byref(lay), ptrs, strides, n_sites, got.ctypes.data, 2) == _cabi.E_DOMAIN


@pytest.mark.parametrize("block, final_newline", [(64, True), (300, False), (5000, True)])
def test_fused_bgzf_read_matches_text_read(tmp_path, block, final_newline):
    """`sai_bgzf_parse_gt` (groups of bgzip blocks inflated and parsed by the same thread) returns
    the rows of the plain-text parse for every group size and thread count -- with blocks much
    shorter than a line (a line spans many blocks and several groups; some groups contain no line
    start at all), a header longer than a group, a last line without a newline, records of other
    chromosomes, the ancestral-allele filter / flip and a region filter."""
    from sai_b200.vcf import _native_read, write_bgzf

    rng = np.random.default_rng(block)
    n_smp, n_rec = 60, 400
    tok = np.array(["0|0", "0|1", "1|1", ".|.", "1/0", "2|1"])
    pos = np.cumsum(rng.integers(1, 40, size=n_rec))
    lines = ["##fileformat=VCFv4.1", "##a long meta line " + "x" * 700,
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"sample_with_a_long_name_{i}" for i in range(n_smp))]
    anc = {}
    for i in range(n_rec):
        chrom = "9" if i % 11 == 3 else "4"
        n_here = n_smp if i % 7 else n_smp  # regular records
        g = rng.choice(len(tok), size=n_here, p=[.7, .1, .08, .04, .04, .04])
        fields = list(tok[g])
        if i % 13 == 0:
            fields[int(rng.integers(n_smp))] = "0|1:35"  # leaves the fast path
        info = "." if i % 5 else "NS=3;DP=14;AF=0.5;" + "k" * int(rng.integers(0, 900))  # lines of very different lengths
        lines.append(f"{chrom}\t{pos[i]}\t.\tA\tG\t.\t.\t{info}\t{'GT:DP' if i % 13 == 0 else 'GT'}\t" + "\t".join(fields))
        if chrom == "4" and rng.random() < 0.8:
            anc[int(pos[i])] = "A" if rng.random() < 0.6 else "G"
    text = ("\n".join(lines) + ("\n" if final_newline else "")).encode()
    plain, bgz = tmp_path / "f.vcf", tmp_path / "f.vcf.gz"
    plain.write_bytes(text)
    write_bgzf(str(bgz), text, block=block)
    req = [(f"sample_with_a_long_name_{i}", int(p)) for i, p in zip(rng.permutation(n_smp)[:25], rng.choice([1, 2, 2, 2, 3], size=25))]
    req2 = [(f"sample_with_a_long_name_{i}", 2) for i in range(10, 50)]  # all diploid, one run
    for requests in (req, req2):
        for anc_table in (None, anc):
            for region in ((None, None), (int(pos[30]), int(pos[300]))):
                p0, g0 = _native_read(str(plain), "4", region[0], region[1], requests, anc_table)
                assert p0.shape[0] > 50
                for kw in (dict(group_blocks=1, n_threads=1), dict(group_blocks=2, n_threads=3), dict(group_blocks=7, n_threads=2), dict(),
                           dict(fused_bgzf=False)):
                    p1, g1 = _native_read(str(bgz), "4", region[0], region[1], requests, anc_table, **kw)
                    assert np.array_equal(p0, p1) and np.array_equal(g0, g1), (block, kw, region, anc_table is not None)


def test_fused_bgzf_read_with_truncated_records(tmp_path):
    """Records with far fewer sample fields than the header promises outnumber the row estimate
    (text size / minimum record size): the fused bgzip read reports that, the reader repeats it with
    more room, and the rows equal the plain-text parse (absent columns = all alleles missing)."""
    from sai_b200.vcf import _native_read, write_bgzf

    n_smp, n_rec = 400, 3000
    lines = ["#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n_smp))]
    for i in range(n_rec):
        lines.append(f"1\t{i + 1}\t.\tA\tG\t.\t.\t.\tGT\t" + "\t".join(["0|1"] * (3 if i % 50 else n_smp)))
    text = ("\n".join(lines) + "\n").encode()
    plain, bgz = tmp_path / "t.vcf", tmp_path / "t.vcf.gz"
    plain.write_bytes(text)
    write_bgzf(str(bgz), text, block=2000)
    req = [(f"s{i}", 2) for i in (0, 2, 3, 399)]
    p0, g0 = _native_read(str(plain), "1", None, None, req, None)
    p1, g1 = _native_read(str(bgz), "1", None, None, req, None)
    assert p0.shape[0] == n_rec and np.array_equal(p0, p1) and np.array_equal(g0, g1)
    assert (g1[1] == [1, 1, -2, -2]).all() and (g1[0] == 1).all()
    # a corrupt block is reported by the fused read as well (payload damaged -> CRC / decoder; ISIZE damaged -> refused)
    import struct

    raw = bytearray(bgz.read_bytes())
    bsize = struct.unpack_from("<H", raw, 16)[0] + 1
    damaged = bytearray(raw)
    damaged[bsize + 40] ^= 0x5A  # inside the second block's deflate data
    (tmp_path / "bad1.vcf.gz").write_bytes(bytes(damaged))
    with pytest.raises(ValueError, match="corrupt"):
        _native_read(str(tmp_path / "bad1.vcf.gz"), "1", None, None, req, None)
    damaged = bytearray(raw)
    bsize2 = struct.unpack_from("<H", raw, bsize + 16)[0] + 1
    struct.pack_into("<I", damaged, bsize + bsize2 - 4, 0x7FFFFFF0)  # ISIZE of the second block
    (tmp_path / "bad2.vcf.gz").write_bytes(bytes(damaged))
    with pytest.raises((ValueError, MemoryError)):
        _native_read(str(tmp_path / "bad2.vcf.gz"), "1", None, None, req, None)


# ---------------------------------------------------------------- BGZF block decoder + CRC-32 (N2)
def test_crc32_matches_zlib():
    """`sai_crc32` (PCLMULQDQ folding where the CPU has it, slicing tables otherwise) == zlib.crc32
    for every length class of the vector path (below 64 bytes, whole 64-byte groups, 16-byte
    groups, ragged tails)."""
    import zlib

    from sai_b200 import _cabi

    lib = _cabi.load()
    rng = np.random.default_rng(0)
    for n in list(range(1, 200)) + [255, 256, 257, 1000, 4096, 65279, 65280, 100003]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for isa in (0, 1):
            assert lib.sai_crc32(d, n, isa) == zlib.crc32(d), (n, isa)


def _deflate_raw(data, level, strategy=None):
    import zlib

    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, zlib.Z_DEFAULT_STRATEGY if strategy is None else strategy)
    return co.compress(data) + co.flush()


def test_inflate_raw_matches_zlib():
    """The in-house raw-deflate decoder of the BGZF reader reproduces zlib's output on dynamic,
    fixed (Z_FIXED) and stored (level 0, incompressible data) blocks, multi-block streams
    (Z_FULL_FLUSH: empty stored blocks in between), VCF-like text, runs with match distances
    1, 2, 3, 4, 8 and long distances, every small size, and never writes past the output."""
    import zlib

    from sai_b200 import _cabi

    lib = _cabi.load()
    rng = np.random.default_rng(1)
    tok = np.array([b"0|0", b"0|1", b"1|0", b"1|1", b".|."])
    vcf_like = b"\n".join(b"1\t%d\t.\tA\tT\t.\tPASS\t.\tGT\t" % (100 + 37 * i) + b"\t".join(tok[(rng.random(400) < 0.05) * rng.integers(1, 5, 400)])
                          for i in range(60))

    def check(data, comp):
        out = np.full(len(data) + 16, 0xCD, dtype=np.uint8)
        assert lib.sai_inflate_raw(comp, len(comp), out.ctypes.data, len(data)) == 1, len(data)
        assert out[: len(data)].tobytes() == data and (out[len(data):] == 0xCD).all()
        # a wrong size is refused, not overrun
        if len(data) > 1:
            assert lib.sai_inflate_raw(comp, len(comp), out.ctypes.data, len(data) - 1) == 0
        assert lib.sai_inflate_raw(comp, len(comp), out.ctypes.data, len(data) + 1) == 0
        assert (out[len(data) + 1:] == 0xCD).all()

    samples = [vcf_like, vcf_like[:65280], bytes(70000), b"ab" * 30000, b"abc" * 20000, b"0|0\t" * 16000, b"abcdefgh" * 8000,
               rng.integers(0, 256, 65280, dtype=np.uint8).tobytes(), rng.integers(0, 4, 50000, dtype=np.uint8).tobytes(),
               (rng.integers(0, 256, 3000, dtype=np.uint8).tobytes() + bytes(29000)) * 2]
    samples += [vcf_like[:n] for n in range(0, 40)] + [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (1, 2, 3, 7, 8, 9, 300)]
    for data in samples:
        for level in (0, 1, 6, 9):
            check(data, _deflate_raw(data, level))
        check(data, _deflate_raw(data, 6, zlib.Z_FIXED))
        check(data, _deflate_raw(data, 6, zlib.Z_HUFFMAN_ONLY))
        check(data, _deflate_raw(data, 6, zlib.Z_RLE))
    # several blocks in one stream, with flush markers (empty stored blocks) between them
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts = [vcf_like[:20000], rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(), bytes(10000), vcf_like[20000:]]
    comp = b"".join(co.compress(p) + co.flush(zlib.Z_FULL_FLUSH if i % 2 else zlib.Z_SYNC_FLUSH) for i, p in enumerate(parts)) + co.flush()
    check(b"".join(parts), comp)


def test_inflate_raw_rejects_corrupt_streams_safely():
    """Bit flips, truncations and random tails: the decoder either refuses the stream or produces
    output that differs from the original (which the block's CRC-32 then catches) -- and in no case
    writes outside the output buffer.  sai_bgzf_inflate reports a corrupt block."""
    import zlib

    from sai_b200 import _cabi
    from sai_b200.vcf import write_bgzf

    lib = _cabi.load()
    rng = np.random.default_rng(2)
    tok = np.array([b"0|0", b"0|1", b"1|1"])
    good = b"\n".join(b"\t".join(tok[(rng.random(500) < 0.1) * rng.integers(1, 3, 500)]) for _ in range(30))
    comp = _deflate_raw(good, 6)
    refused = 0
    for it in range(1500):
        bad = bytearray(comp)
        kind = it % 3
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            bad = bad[: int(rng.integers(0, len(bad)))]
        else:
            at = int(rng.integers(0, len(bad)))
            bad[at:] = rng.integers(0, 256, int(rng.integers(1, 60)), dtype=np.uint8).tobytes()
        out = np.full(len(good) + 32, 0xCD, dtype=np.uint8)
        ok = lib.sai_inflate_raw(bytes(bad), len(bad), out.ctypes.data, len(good))
        assert (out[len(good):] == 0xCD).all()
        if ok and out[: len(good)].tobytes() != good:
            assert lib.sai_crc32(out.ctypes.data, len(good), 0) != zlib.crc32(good)
        refused += not ok
    assert refused > 1000


def test_plain_gzip_fast_path(tmp_path):
    """`sai_gzip_inflate` (a single-member gzip file in one native call) returns gzip's text for
    headers with and without a file name, refuses files with several members or a damaged payload
    (SAI_E_DOMAIN: the reader then streams through Python's gzip module), and `_native_read` gives
    the same rows for a plain file, its one-member gzip and a two-member gzip."""
    import gzip

    from sai_b200 import _cabi
    from sai_b200.vcf import _native_read

    lib = _cabi.load()
    rng = np.random.default_rng(4)
    n_smp, n_rec = 30, 500
    tok = np.array(["0|0", "0|1", "1|1", ".|1"])
    lines = ["##x", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n_smp))]
    for i in range(n_rec):
        lines.append(f"5\t{3 * i + 1}\t.\tC\tT\t.\t.\t.\tGT\t" + "\t".join(tok[rng.integers(0, 4, n_smp)]))
    text = ("\n".join(lines) + "\n").encode()
    plain, one, named, two = tmp_path / "p.vcf", tmp_path / "one.vcf.gz", tmp_path / "named.vcf.gz", tmp_path / "two.vcf.gz"
    plain.write_bytes(text)
    one.write_bytes(gzip.compress(text, 6))
    with gzip.GzipFile(filename="inner_name.vcf", mode="wb", fileobj=open(named, "wb")) as f:
        f.write(text)
    half = text.index(b"\n", len(text) // 2) + 1
    two.write_bytes(gzip.compress(text[:half]) + gzip.compress(text[half:]))
    for path, want in ((one, len(text)), (named, len(text)), (two, _cabi.E_DOMAIN)):
        raw = np.frombuffer(path.read_bytes(), dtype=np.uint8)
        out = np.full(len(text) + 8, 0xCD, dtype=np.uint8)
        got = lib.sai_gzip_inflate(raw.ctypes.data, raw.size, out.ctypes.data, len(text))
        assert got == want, path
        if want > 0:
            assert out[: len(text)].tobytes() == text and (out[len(text):] == 0xCD).all()
            assert lib.sai_gzip_inflate(raw.ctypes.data, raw.size, out.ctypes.data, len(text) - 1) == _cabi.E_CAPACITY
            bad = raw.copy()
            bad[raw.size // 2] ^= 0x10
            assert lib.sai_gzip_inflate(bad.ctypes.data, bad.size, out.ctypes.data, len(text)) == _cabi.E_DOMAIN
    req = [(f"s{i}", 2) for i in (3, 4, 5, 29, 0)]
    p0, g0 = _native_read(str(plain), "5", None, None, req, None)
    assert p0.shape[0] == n_rec
    for path in (one, named, two):
        p1, g1 = _native_read(str(path), "5", None, None, req, None)
        assert np.array_equal(p0, p1) and np.array_equal(g0, g1), path


# ---------------------------------------------------------------- whole-genome sharding (host logic)
def test_shard_genome_follows_split_windows_ranges():
    """`shard_genome` cuts the flattened (chromosome, window) list like
    ChunkGenerator._split_windows_ranges (chunk_generator.py:130-141): on one chromosome the pieces
    ARE its ranges (incl. the reference's own KAT [(1, 30000), (25001, 55000)],
    tests/generators/test_chunk_generator.py:39); across chromosomes every window lands in exactly
    one piece, pieces never straddle a chromosome, shard sizes differ by at most one window, and a
    piece's site range is the region read `chr:first.start-last.end` (the halo)."""
    import sai_oracle as orc
    from sai_b200.genome import Piece, piece_site_range, shard_genome
    from sai_b200.windows import split_genome, split_windows_ranges

    wins = split_genome([1, 46000], 30000, 5000)  # the reference KAT's grid: window_size 30000, step 5000... checked below
    two = shard_genome([wins], 2)
    ranges = [(wins[p[0].win_lo][0], wins[p[0].win_hi - 1][1]) for p in two]
    assert ranges == split_windows_ranges(wins, 2) == orc.split_windows_ranges(wins, 2)
    kat = split_genome([1, 30000], 30000, 25000)  # [(1, 30000), (25001, 55000)]
    assert [(kat[p[0].win_lo][0], kat[p[0].win_hi - 1][1]) for p in shard_genome([kat], 2)] == [(1, 30000), (25001, 55000)]

    rng = np.random.default_rng(4)
    chrom_pos = [np.cumsum(rng.integers(1, 400, size=n)).astype(np.int32) for n in (900, 40, 333, 1, 700)]
    chrom_wins = [split_genome([int(p[0]), int(p[-1])], 20000, 5000) for p in chrom_pos]
    total = sum(len(w) for w in chrom_wins)
    for world in (1, 2, 3, 8, 64, total + 5):
        shards = shard_genome(chrom_wins, world)
        assert len(shards) == world
        sizes = [sum(p.win_hi - p.win_lo for p in s) for s in shards]
        assert sum(sizes) == total and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
        seen = []
        for s in shards:
            for p in s:
                assert 0 <= p.win_lo < p.win_hi <= len(chrom_wins[p.chrom])
                seen += [(p.chrom, i) for i in range(p.win_lo, p.win_hi)]
        assert seen == [(c, i) for c, w in enumerate(chrom_wins) for i in range(len(w))]  # genome order, once each
        for s in shards:
            for p in s:
                pos, w = chrom_pos[p.chrom], chrom_wins[p.chrom]
                lo, hi = piece_site_range(pos, w, p, align=False)
                inside = (pos >= w[p.win_lo][0]) & (pos <= w[p.win_hi - 1][1])
                assert (hi - lo) == int(inside.sum()) and (hi == lo or (inside[lo] and inside[hi - 1]))
                lo_a, hi_a = piece_site_range(pos, w, p)
                assert lo_a % 32 == 0 and lo - 31 <= lo_a <= lo and hi_a == max(hi, lo_a)
    assert shard_genome([[]], 3) == [[], [], []]


@pytest.mark.parametrize("layout", ["contiguous", "scattered"])
def test_all_diploid_requests_match_python_reader(tmp_path, layout):
    """When every request is diploid the native parser takes the allele SUMS of a regular record in
    one vector sweep (flipped records: |a - 1| per allele inside the sweep) and moves them either
    as one block per run of neighbouring sample columns ("contiguous") or column by column
    ("scattered": permuted and duplicated columns).  Both == the pure-Python reader, with
    and without the ancestral-allele table, whole file and region, "." and multi-allelic alleles,
    irregular records in between (field walker)."""
    from sai_b200.configs import PloidyConfig
    from sai_b200.vcf import read_data

    rng = np.random.default_rng(len(layout))
    n_samples, n_sites = 203, 700
    lines = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n_samples))]
    pos = np.cumsum(rng.integers(1, 50, size=n_sites))
    bases = "ACGT"
    recs = []
    for p in pos:
        ref = bases[rng.integers(4)]
        alt = bases[(bases.index(ref) + 1 + rng.integers(3)) % 4]
        toks = []
        for _ in range(n_samples):
            a = [("." if rng.random() < 0.05 else str(int(rng.integers(0, 10 if rng.random() < 0.03 else 2)))) for _ in range(2)]
            toks.append(a[0] + ("|" if rng.random() < 0.7 else "/") + a[1])
        kind = rng.random()
        if kind < 0.05:
            toks[int(rng.integers(n_samples))] = "1"  # haploid field: the record leaves the fast path
        lines.append("\t".join(["7", str(p), ".", ref, alt, ".", "PASS", ".", "GT"] + toks))
        recs.append((int(p), ref, alt))
    vcf = tmp_path / "d.vcf"
    vcf.write_text("\n".join(lines) + "\n")
    anc = tmp_path / "anc.bed"
    with open(anc, "w") as f:
        for p, ref, alt in recs:
            r = rng.random()
            if r >= 0.1:
                f.write(f"7\t{p - 1}\t{p}\t{ref if r < 0.5 else (alt if r < 0.92 else 'N')}\n")
    if layout == "contiguous":
        groups = {"ref": list(range(0, 120)), "tgt": list(range(120, 200)), "src": [200, 201, 202]}
    else:
        perm = rng.permutation(n_samples)
        groups = {"ref": [int(x) for x in perm[:90]], "tgt": [int(x) for x in perm[60:140]], "src": [int(perm[0]), 202, 0]}
    for g, idx in groups.items():
        (tmp_path / f"{g}.list").write_text("".join(f"{g.upper()}\ts{i}\n" for i in idx))
    pc = PloidyConfig({"ref": {"REF": 2}, "tgt": {"TGT": 2}, "src": {"SRC": 2}})
    lists = [str(tmp_path / f"{g}.list") for g in ("ref", "tgt", "src")]
    for anc_file in (None, str(anc)):
        for region in ((None, None), (int(pos[40]), int(pos[600]))):
            a = read_data(str(vcf), "7", pc, *lists, None, anc_file, start=region[0], end=region[1], native=True)
            b = read_data(str(vcf), "7", pc, *lists, None, anc_file, start=region[0], end=region[1], native=False)
            for g in ("ref", "tgt", "src"):
                assert list(a[g][0]) == list(b[g][0])
                for p in a[g][0]:
                    assert np.array_equal(a[g][0][p].POS, b[g][0][p].POS), (g, p)
                    assert np.array_equal(a[g][0][p].GT, b[g][0][p].GT), (g, p, anc_file, region)
                    assert a[g][0][p].POS.size > 300


def test_native_parser_absent_columns_and_runs():
    """A record with fewer sample fields than the header promises: the missing columns read as all
    alleles missing (-ploidy), in the block-copy path (runs of neighbouring columns, all diploid), the
    column-by-column path and the mixed-ploidy path alike; and the three paths agree on every value."""
    import ctypes as C

    from sai_b200 import _cabi

    lib = _cabi.load()
    rng = np.random.default_rng(3)
    n_smp, n_rec = 90, 40
    tok = np.array(["0|0", "0|1", "1/1", ".|1", "2|0"])
    lines = ["#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n_smp))]
    n_fields = []
    for i in range(n_rec):
        nf = n_smp if i % 3 else int(rng.integers(1, n_smp))
        n_fields.append(nf)
        lines.append(f"1\t{10 * i + 5}\t.\tA\tG\t.\t.\t.\tGT\t" + "\t".join(tok[rng.integers(0, len(tok), size=nf)]))
    text = ("\n".join(lines) + "\n").encode()

    def parse(cols, ploidy):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        pl = np.ascontiguousarray(ploidy, dtype=np.int32)
        out_pos = np.empty(n_rec, dtype=np.int32)
        out_gt = np.full((n_rec, cols.size + 3), 99, dtype=np.int8)  # row stride > n_out
        consumed = C.c_int64(0)
        n = lib.sai_vcf_parse_gt(C.c_char_p(text), len(text), b"1", 1, 0, cols.ctypes.data, pl.ctypes.data, cols.size, None, None, 0,
                                 out_pos.ctypes.data, out_gt.ctypes.data, out_gt.shape[1], n_rec, C.byref(consumed), 2)
        assert n == n_rec and consumed.value == len(text) and (out_gt[:, cols.size:] == 99).all()
        return out_gt[:, : cols.size].copy()

    everything = np.arange(n_smp)
    blocks = parse(everything, [2] * n_smp)  # one run of 90 columns: block copies
    for i, nf in enumerate(n_fields):
        assert (blocks[i, nf:] == -2).all() and (blocks[i, :nf] >= -2).all() and (blocks[i, :nf] != -2).any()
    perm = rng.permutation(n_smp)
    assert np.array_equal(parse(perm, [2] * n_smp), blocks[:, perm])  # column by column
    mixed = parse(np.concatenate([everything, [0]]), [2] * n_smp + [1])  # one haploid request: the a0 / a1 path
    assert np.array_equal(mixed[:, :n_smp], blocks)


@pytest.mark.parametrize("crlf, final_newline", [(False, True), (True, True), (False, False)])
def test_regular_record_fast_path_matches_python_reader(tmp_path, crlf, final_newline):
    """Records whose sample fields are all `x|y` / `x/y` take the vector fast path of the native
    parser (16 fields per AVX-512 step); records with anything else in them (a two-digit allele, a
    haploid or triploid field, GT:DP) fall back to the field walker.  Both must give exactly what
    the pure-Python reader gives: "." alleles, multi-allelic indices, requested ploidy 1 / 2 / 3 / 4
    on diploid fields (cut or padded with missing), flipped records (|a - 1| on every allele, the
    padded ones too), a sample requested by two populations, region reads."""
    from sai_b200.configs import PloidyConfig
    from sai_b200.vcf import read_data

    rng = np.random.default_rng(5 + crlf + 2 * final_newline)
    n_samples, n_sites = 131, 600  # 131 fields: eight full 16-field steps and a ragged tail
    nl = "\r\n" if crlf else "\n"
    lines = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(n_samples))]
    pos = np.cumsum(rng.integers(1, 50, size=n_sites))
    bases = "ACGT"
    recs, n_irregular = [], 0
    for p in pos:
        ref = bases[rng.integers(4)]
        alt = bases[(bases.index(ref) + 1 + rng.integers(3)) % 4]
        toks = []
        for _ in range(n_samples):
            a = [("." if rng.random() < 0.05 else str(int(rng.integers(0, 10 if rng.random() < 0.03 else 2)))) for _ in range(2)]
            toks.append(a[0] + ("|" if rng.random() < 0.7 else "/") + a[1])
        fmt = "GT"
        kind = rng.random()
        if kind < 0.04:
            toks[int(rng.integers(n_samples))] = "10|1"  # two-digit allele
        elif kind < 0.08:
            toks[int(rng.integers(n_samples))] = "1"  # haploid field
        elif kind < 0.12:
            toks[int(rng.integers(n_samples))] = "0/1/1"
        elif kind < 0.16:
            fmt, toks = "GT:DP", [t + ":7" for t in toks]
        n_irregular += kind < 0.16
        lines.append("\t".join(["3", str(p), ".", ref, alt, ".", "PASS", ".", fmt] + toks))
        recs.append((int(p), ref, alt))
    assert 20 < n_irregular < 200
    vcf = tmp_path / "reg.vcf"
    open(vcf, "w", newline="").write(nl.join(lines) + (nl if final_newline else ""))
    anc = tmp_path / "anc.bed"
    with open(anc, "w") as f:
        for p, ref, alt in recs:
            r = rng.random()
            if r < 0.1:
                continue
            f.write(f"3\t{p - 1}\t{p}\t{ref if r < 0.5 else (alt if r < 0.92 else 'N')}\n")
    (tmp_path / "ref.list").write_text("".join(f"R1\ts{i}\n" for i in range(0, 70)) + "".join(f"R4\ts{i}\n" for i in range(60, 100)))
    (tmp_path / "tgt.list").write_text("".join(f"T1\ts{i}\n" for i in range(90, 131)) + "T3\ts5\nT3\ts130\n")
    (tmp_path / "src.list").write_text("S\ts0\nS\ts64\nS\ts129\n")
    pc = PloidyConfig({"ref": {"R1": 2, "R4": 4}, "tgt": {"T1": 1, "T3": 3}, "src": {"S": 2}})
    lists = [str(tmp_path / f"{g}.list") for g in ("ref", "tgt", "src")]
    for anc_file in (None, str(anc)):
        for region in ((None, None), (int(pos[50]), int(pos[500]))):
            a = read_data(str(vcf), "3", pc, *lists, None, anc_file, start=region[0], end=region[1], native=True)
            b = read_data(str(vcf), "3", pc, *lists, None, anc_file, start=region[0], end=region[1], native=False)
            rows = 0
            for g in ("ref", "tgt", "src"):
                assert list(a[g][0]) == list(b[g][0])
                for p in a[g][0]:
                    assert np.array_equal(a[g][0][p].POS, b[g][0][p].POS), (g, p)
                    assert np.array_equal(a[g][0][p].GT, b[g][0][p].GT), (g, p, anc_file, region)
                    rows = a[g][0][p].POS.size
            assert rows > 300
