"""Host-side encoding: per-individual allele-sum matrices -> tiled bit-planes.

Replaces the reference's in-memory representation -- one int64 matrix
``(sites, individuals)`` per population produced by
``reshape_genotypes(is_phased=False)`` (sai/utils/utils.py:405-410) -- by the
packed layout of ``include/sai_b200.h`` (2 bits per individual for ploidy <= 2).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _cabi


class PopData:
    """Positions + per-individual allele sums of one population (the two
    fields of the reference's ``ChromosomeData`` the scoring path reads,
    sai/utils/genomic_dataclasses.py:25-46)."""

    __slots__ = ("POS", "GT")

    def __init__(self, POS, GT):
        self.POS = np.asarray(POS)
        self.GT = np.asarray(GT)


def make_layout(n_samples: Sequence[int], ploidy: Sequence[int], bits: Optional[Sequence[int]] = None):
    lib = _cabi.load()
    n = len(n_samples)
    if n != len(ploidy):
        raise ValueError("n_samples and ploidy must have the same length")
    if n > _cabi.MAX_POPS:
        raise ValueError(f"at most {_cabi.MAX_POPS} populations per packed matrix")
    for p in ploidy:
        if not isinstance(p, (int, np.integer)) or isinstance(p, bool) or p <= 0:
            raise ValueError("ploidy must be a positive integer.")
    lay = _cabi.Layout()
    ns = (C.c_int32 * n)(*[int(x) for x in n_samples])
    pl = (C.c_int32 * n)(*[int(x) for x in ploidy])
    bt = (C.c_int32 * n)(*[int(x) for x in (bits if bits is not None else [0] * n)])
    _cabi.check(lib.sai_layout_init(C.byref(lay), n, ns, pl, bt))
    return lay


@dataclass
class PackedGenotypes:
    layout: "_cabi.Layout"
    n_sites: int
    pos: np.ndarray  # int32, sorted, unique
    packed: np.ndarray  # uint8, host
    pop_names: list = field(default_factory=list)
    # negative-value table (DD only; include/sai_b200.h "N4"): raw values of the calls the
    # bit-planes code as missing, per population sorted by (site, individual)
    neg_off: Optional[np.ndarray] = None  # int64 [n_pops + 1]
    neg_site: Optional[np.ndarray] = None  # int32
    neg_ind: Optional[np.ndarray] = None  # int32
    neg_val: Optional[np.ndarray] = None  # int32

    @property
    def n_tiles(self) -> int:
        return (self.n_sites + _cabi.TILE_SITES - 1) // _cabi.TILE_SITES

    @property
    def nbytes(self) -> int:
        return int(self.packed.nbytes)


@dataclass
class MatrixGenotypes:
    """Unpacked input of ``HostEngine.score_matrices``: one int8 matrix ``(sites, individuals)``
    per population (row-strided views of one parsed matrix are fine) plus the layout the engine
    packs them into on the fly."""

    layout: "_cabi.Layout"
    n_sites: int
    pos: np.ndarray
    mats: list
    pop_names: list = field(default_factory=list)
    neg_off: Optional[np.ndarray] = None
    neg_site: Optional[np.ndarray] = None
    neg_ind: Optional[np.ndarray] = None
    neg_val: Optional[np.ndarray] = None


def _as_i8(gt: np.ndarray) -> np.ndarray:
    gt = np.asarray(gt)
    if gt.ndim != 2:
        raise ValueError("genotype matrix must be 2-D (sites x individuals)")
    if gt.dtype == np.int8 and (gt.flags.c_contiguous or (gt.strides[1] == 1 and gt.strides[0] >= gt.shape[1])):
        return gt  # row-strided views (column blocks of one parsed matrix) are packed in place
    # negative = missing (any negative value, sai/stats/stat_utils.py:45).  8 bit-planes hold every
    # non-negative int8; a per-individual allele sum above 127 has no encoding
    if gt.size and int(gt.max()) > 127:
        raise ValueError("per-individual allele sums above 127 are not supported by the packed layout")
    return np.ascontiguousarray(np.maximum(gt, -1).astype(np.int8))


def bits_for(gt: np.ndarray, ploidy: int) -> int:
    vmax = int(gt.max()) if gt.size else 0
    return int(_cabi.load().sai_bits_for_max_value(max(int(ploidy), vmax)))


def pack_populations(
    gts: Sequence[np.ndarray],
    ploidy: Sequence[int],
    pos: np.ndarray,
    pop_names: Optional[Sequence[str]] = None,
    bits: Optional[Sequence[int]] = None,
    n_threads: int = 0,
    out: Optional[np.ndarray] = None,
    keep_negatives: bool = False,
) -> PackedGenotypes:
    """Packs one matrix per population (all over the same ``pos``).
    ``keep_negatives`` also records the raw values of the missing calls (any
    ``v < 0``) in the negative-value table the DD statistic needs."""
    lib = _cabi.load()
    neg = negative_table(gts) if keep_negatives else (None, None, None, None)
    mats = [_as_i8(g) for g in gts]
    pos = np.ascontiguousarray(np.asarray(pos), dtype=np.int32)
    n_sites = int(pos.shape[0])
    for m in mats:
        if m.shape[0] != n_sites:
            raise ValueError("every population must have one row per position")
    if bits is None:
        bits = [bits_for(m, p) for m, p in zip(mats, ploidy)]
    lay = make_layout([m.shape[1] for m in mats], ploidy, bits)
    nbytes = int(lib.sai_packed_bytes(C.byref(lay), n_sites))
    if out is None:
        out = np.empty(max(nbytes, 1), dtype=np.uint8)
    elif out.nbytes < nbytes or out.dtype != np.uint8:
        raise ValueError("output buffer too small")
    ptrs = (C.c_void_p * len(mats))(*[m.ctypes.data for m in mats])
    strides = (C.c_int64 * len(mats))(*[m.strides[0] if n_sites > 1 else max(1, m.shape[1]) for m in mats])
    _cabi.check(lib.sai_pack_i8_all(C.byref(lay), ptrs, strides, n_sites, out.ctypes.data, n_threads))
    return PackedGenotypes(lay, n_sites, pos, out[:nbytes] if nbytes else out[:0], list(pop_names or []), *neg)


def negative_table(gts: Sequence[np.ndarray]):
    """``(neg_off, neg_site, neg_ind, neg_val)`` of a list of per-population
    matrices: every entry ``v < 0`` (a missing call; ``-1`` for ``0/.``, ``-2`` for
    ``./.`` in diploid data, sai/utils/utils.py:405-410), row-major, i.e. sorted by
    (site, individual) inside a population."""
    lib = _cabi.load()
    off = np.zeros(len(gts) + 1, dtype=np.int64)
    sites, inds, vals = [], [], []
    for i, g in enumerate(gts):
        g = np.asarray(g)
        if g.dtype == np.int8 and g.ndim == 2 and g.shape[1] >= 1 and g.strides[1] == 1 and g.strides[0] >= g.shape[1]:
            # native two-pass scan (multi-threaded); int8 is what the VCF reader produces
            stride = g.strides[0] if g.shape[0] > 1 else g.shape[1]
            n = int(lib.sai_neg_table_i8(g.ctypes.data, g.shape[0], g.shape[1], stride, None, None, None, 0, 0))
            if n < 0:
                _cabi.check(n)
            r, c, v = (np.empty(n, dtype=np.int32) for _ in range(3))
            if n:
                m = int(lib.sai_neg_table_i8(g.ctypes.data, g.shape[0], g.shape[1], stride, r.ctypes.data,
                                             c.ctypes.data, v.ctypes.data, n, 0))
                assert m == n
            sites.append(r)
            inds.append(c)
            vals.append(v)
        else:
            r, c = np.nonzero(g < 0)
            sites.append(r.astype(np.int32))
            inds.append(c.astype(np.int32))
            vals.append(np.maximum(g[r, c].astype(np.int64), -(2**31) + 1).astype(np.int32))
        off[i + 1] = off[i] + sites[-1].shape[0]
    cat = lambda parts: np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0, np.int32)
    return off, cat(sites), cat(inds), cat(vals)


def unpack_population(pg: PackedGenotypes, pop: int, site0: int = 0, n: Optional[int] = None) -> np.ndarray:
    """Decodes population ``pop`` back to int8 allele sums (missing = -1)."""
    lib = _cabi.load()
    if n is None:
        n = pg.n_sites - site0
    ns = pg.layout.pop[pop].n_samples
    out = np.empty((n, ns), dtype=np.int8)
    if n:
        _cabi.check(
            lib.sai_unpack_i8(
                C.byref(pg.layout), pop, pg.packed.ctypes.data, pg.n_sites, site0, n, out.ctypes.data, ns
            )
        )
    return out


@dataclass
class ZtGenotypes:
    """A packed matrix in zt form (zero-suppressed tiles, include/sai_b200.h "zt"):
    what goes over PCIe.  Lossless; the device rebuilds the dense tiles."""

    layout: "_cabi.Layout"
    n_sites: int
    pos: np.ndarray
    stream: np.ndarray  # uint8, host
    tile_off: np.ndarray  # uint64 [n_tiles + 1]
    pop_names: list = field(default_factory=list)
    neg_off: Optional[np.ndarray] = None
    neg_site: Optional[np.ndarray] = None
    neg_ind: Optional[np.ndarray] = None
    neg_val: Optional[np.ndarray] = None

    @property
    def n_tiles(self) -> int:
        return (self.n_sites + _cabi.TILE_SITES - 1) // _cabi.TILE_SITES

    @property
    def nbytes(self) -> int:
        return int(self.stream.nbytes) + int(self.tile_off.nbytes)


def compress(pg: PackedGenotypes, n_threads: int = 0, out: Optional[np.ndarray] = None) -> ZtGenotypes:
    """Packed tiles -> zt stream (host, multi-threaded).  ``out``: optional uint8
    buffer (e.g. pinned memory) of at least ``sai_zt_bound`` bytes."""
    lib = _cabi.load()
    n_tiles = pg.n_tiles
    tile_off = np.zeros(n_tiles + 1, dtype=np.uint64)
    bound = int(lib.sai_zt_bound(C.byref(pg.layout), pg.n_sites))
    buf = np.empty(max(bound, 8), dtype=np.uint8) if out is None else out
    if buf.dtype != np.uint8 or not buf.flags.c_contiguous:
        raise ValueError("out must be a contiguous uint8 buffer")
    n = lib.sai_zt_encode(C.byref(pg.layout), pg.packed.ctypes.data if pg.packed.size else None, pg.n_sites,
                          buf.ctypes.data, buf.nbytes, tile_off.ctypes.data, n_threads)
    if n < 0:
        _cabi.check(int(n))
    stream = buf[: int(n)] if out is not None else buf[: int(n)].copy()
    return ZtGenotypes(pg.layout, pg.n_sites, pg.pos, stream, tile_off, list(pg.pop_names),
                       pg.neg_off, pg.neg_site, pg.neg_ind, pg.neg_val)


def compress_matrices(mg: MatrixGenotypes, n_threads: int = 0, out: Optional[np.ndarray] = None) -> ZtGenotypes:
    """int8 matrices -> zt stream in one pass (``sai_zt_pack_i8``): every tile is packed and
    encoded while it sits in a packer thread's L1, the dense tiles never reach memory.  Same
    bytes as ``compress(pack_populations(...))``.  Raises ``ValueError`` when a value does not
    fit ``mg.layout``'s bit-planes."""
    lib = _cabi.load()
    n_pops = mg.layout.n_pops
    mats = [_as_i8(m) for m in mg.mats]
    if len(mats) != n_pops or any(m.shape[0] != mg.n_sites for m in mats):
        raise ValueError("one (n_sites x individuals) matrix per population of the layout")
    n_tiles = (mg.n_sites + _cabi.TILE_SITES - 1) // _cabi.TILE_SITES
    tile_off = np.zeros(n_tiles + 1, dtype=np.uint64)
    bound = int(lib.sai_zt_bound(C.byref(mg.layout), mg.n_sites))
    buf = np.empty(max(bound, 8), dtype=np.uint8) if out is None else out
    if buf.dtype != np.uint8 or not buf.flags.c_contiguous:
        raise ValueError("out must be a contiguous uint8 buffer")
    ptrs = (C.c_void_p * n_pops)(*[m.ctypes.data for m in mats])
    strides = (C.c_int64 * n_pops)(*[m.strides[0] if m.shape[0] > 1 else max(m.shape[1], 1) for m in mats])
    n = lib.sai_zt_pack_i8(C.byref(mg.layout), ptrs, strides, mg.n_sites, buf.ctypes.data, buf.nbytes, tile_off.ctypes.data, n_threads)
    if n < 0:
        _cabi.check(int(n))  # E_DOMAIN -> ValueError
    stream = buf[: int(n)] if out is not None else buf[: int(n)].copy()
    return ZtGenotypes(mg.layout, mg.n_sites, mg.pos, stream, tile_off, list(mg.pop_names),
                       mg.neg_off, mg.neg_site, mg.neg_ind, mg.neg_val)


def decompress(zt: ZtGenotypes) -> PackedGenotypes:
    """Host decoder (tests / tools): zt stream -> packed tiles."""
    lib = _cabi.load()
    nbytes = int(lib.sai_packed_bytes(C.byref(zt.layout), zt.n_sites))
    packed = np.empty(max(nbytes, 1), dtype=np.uint8)
    _cabi.check(lib.sai_zt_decode_host(C.byref(zt.layout), zt.stream.ctypes.data if zt.stream.size else None,
                                       zt.tile_off.ctypes.data, zt.n_sites, packed.ctypes.data))
    return PackedGenotypes(zt.layout, zt.n_sites, zt.pos, packed[:nbytes], list(zt.pop_names),
                           zt.neg_off, zt.neg_site, zt.neg_ind, zt.neg_val)
