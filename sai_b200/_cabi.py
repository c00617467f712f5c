"""ctypes binding of ``include/sai_b200.h`` (the C-ABI drop-in boundary).

The product path has no CPU fallback: if ``libsai_b200.so`` is missing it is
built with nvcc; if that fails, importing this module raises.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import build as _build

MAX_POPS = 16
MAX_SRC = 8
MAX_JOBS = 8
TILE_SITES = 32

OK, E_ARG, E_CUDA, E_DOMAIN, E_CAPACITY, E_NOMEM = 0, -1, -2, -3, -4, -5
OPS = {"=": 0, "<": 1, ">": 2, "<=": 3, ">=": 4}


class PopLayout(C.Structure):
    _fields_ = [
        ("n_samples", C.c_int32),
        ("ploidy", C.c_int32),
        ("bits", C.c_int32),
        ("pair_off", C.c_int32),
        ("n_pairs", C.c_int32),
        ("n_groups", C.c_int32),
    ]


class Layout(C.Structure):
    _fields_ = [
        ("n_pops", C.c_int32),
        ("pairs_per_site", C.c_int32),
        ("pop", PopLayout * MAX_POPS),
    ]


class Cond(C.Structure):
    _fields_ = [
        ("w", C.c_double),
        ("y", C.c_double * MAX_SRC),
        ("one_minus_y", C.c_double * MAX_SRC),
        ("op", C.c_int32 * MAX_SRC),
        ("enabled", C.c_int32),
        ("pad_", C.c_int32),
    ]


class Job(C.Structure):
    _fields_ = [
        ("ref_pop", C.c_int32),
        ("tgt_pop", C.c_int32),
        ("n_src", C.c_int32),
        ("src_pop", C.c_int32 * MAX_SRC),
        ("anc_allele_available", C.c_int32),
        ("u", Cond),
        ("x", C.c_double),
        ("q", Cond),
        ("quantile", C.c_double),
    ]


class HostResults(C.Structure):
    _fields_ = [
        ("nsnps", C.c_void_p),
        ("u", C.c_void_p),
        ("q", C.c_void_p),
        ("q_cnt", C.c_void_p),
        ("u_start", C.c_void_p),
        ("q_start", C.c_void_p),
        ("totals", C.c_void_p),
        ("u_cand", C.c_void_p),
        ("q_cand", C.c_void_p),
        ("cap_u", C.c_int64),
        ("cap_q", C.c_int64),
    ]


#: every symbol include/sai_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_I32, _I64, _U64 = C.c_int32, C.c_int64, C.c_uint64
_LAY = C.POINTER(Layout)
_JOB = C.POINTER(Job)
SYMBOLS = {
    "sai_version": (C.c_char_p, []),
    "sai_last_error": (C.c_char_p, []),
    "sai_layout_init": (C.c_int, [_LAY, _I32, _P, _P, _P]),
    "sai_bits_for_max_value": (_I32, [_I32]),
    "sai_num_tiles": (_I64, [_I64]),
    "sai_packed_bytes": (_U64, [_LAY, _I64]),
    "sai_pack_i8": (C.c_int, [_LAY, _I32, _P, _I64, _I64, _P, _I32]),
    "sai_pack_i8_all": (C.c_int, [_LAY, _P, _P, _I64, _P, _I32]),
    "sai_pack_isa": (C.c_char_p, []),
    "sai_pack_i8_isa": (C.c_int, [_LAY, _I32, _P, _I64, _I64, _P, _I32, _I32]),
    "sai_unpack_i8": (C.c_int, [_LAY, _I32, _P, _I64, _I64, _I64, _P, _I64]),
    "sai_neg_table_i8": (_I64, [_P, _I64, _I32, _I64, _P, _P, _P, _I64, _I32]),
    "sai_vcf_parse_gt": (
        _I64,
        [_P, _I64, C.c_char_p, _I64, _I64, _P, _P, _I32, _P, _P, _I64, _P, _P, _I64, _I64, C.POINTER(_I64), _I32],
    ),
    "sai_vcf_chrom_span": (C.c_int, [_P, _I64, C.c_char_p, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64), _I32]),
    "sai_is_bgzf": (_I32, [_P, _I64]),
    "sai_bgzf_scan": (_I64, [_P, _I64, _I64, _I64, _P, _P, C.POINTER(_I64)]),
    "sai_bgzf_inflate": (C.c_int, [_P, _P, _P, _I64, _P, _I32]),
    "sai_bgzf_parse_gt": (_I64, [_P, _P, _P, _I64, _I64, C.c_char_p, _I64, _I64, _P, _P, _I32, _P, _P, _I64, _P, _P, _I64, _I64, _I32, _I32]),
    "sai_gzip_inflate": (_I64, [_P, _I64, _P, _I64]),
    "sai_inflate_raw": (_I32, [_P, _I64, _P, _I64]),
    "sai_crc32": (C.c_uint32, [_P, _I64, _I32]),
    "sai_site_counts": (C.c_int, [_LAY, _P, _I64, _I64, _P, _P, _I64, _I32, _P]),
    "sai_site_flags": (
        C.c_int,
        [_LAY, _P, _I64, _I64, _I64, _JOB, _I32, _P, _P, _P, _I64, _P, _P, _I64, _I32, _P],
    ),
    "sai_flags_from_counts": (C.c_int, [_LAY, _P, _P, _I64, _I64, _JOB, _I32, _P, _P, _P, _I64, _P]),
    "sai_window_stats": (
        C.c_int,
        [_P, _I64, _P, _P, _I64, _JOB, _I32, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _I64, _P],
    ),
    "sai_window_stats_pieces": (
        C.c_int,
        [_P, _I64, _P, _P, _P, _P, _I64, _JOB, _I32, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _I64, _P],
    ),
    "sai_column_quantiles": (C.c_int, [_P, _I32, _I32, _I64, _I64, _I64, C.c_double, _P, _P]),
    "sai_window_patterns": (
        C.c_int,
        [_LAY, _P, _I64, _P, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _P, _I32, _P, _P],
    ),
    "sai_engine_pattern_sums": (C.c_int, [_P, _LAY, _I32, _I32, _I32, _P, _I32, _P]),
    "sai_hist_rows": (_I64, [_LAY, _P, _I32]),
    "sai_site_hist": (C.c_int, [_LAY, _P, _I64, _P, _I32, _P, _I64, _P, _P]),
    "sai_window_dd": (
        C.c_int,
        [_LAY, _P, _P, _I64, _P, _P, _I64, _P, _I64, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _I32, _P, _P],
    ),
    "sai_engine_dd_sums": (C.c_int, [_P, _LAY, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _I32]),
    "sai_engine_create": (C.c_int, [_I32, C.POINTER(_P)]),
    "sai_engine_destroy": (None, [_P]),
    "sai_engine_score_host": (
        C.c_int,
        [_P, _LAY, _P, _P, _I64, _P, _P, _I64, _JOB, _I32, C.POINTER(HostResults)],
    ),
    "sai_engine_score_host_zt": (
        C.c_int,
        [_P, _LAY, _P, _P, _P, _I64, _P, _P, _I64, _JOB, _I32, C.POINTER(HostResults)],
    ),
    "sai_engine_score_host_i8": (
        C.c_int,
        [_P, _LAY, _P, _P, _P, _I64, _P, _P, _I64, _JOB, _I32, C.POINTER(HostResults)],
    ),
    "sai_engine_set_host_threads": (C.c_int, [_P, _I32]),
    "sai_engine_set_i8_wire": (C.c_int, [_P, _I32]),
    "sai_engine_i8_wire_bytes": (_U64, [_P]),
    "sai_engine_score_resident": (C.c_int, [_P, _JOB, _I32, C.POINTER(HostResults)]),
    "sai_engine_rescore_windows": (C.c_int, [_P, C.POINTER(HostResults)]),
    "sai_zt_bound": (_U64, [_LAY, _I64]),
    "sai_zt_encode": (_I64, [_LAY, _P, _I64, _P, _U64, _P, _I32]),
    "sai_zt_encode_isa": (_I64, [_LAY, _P, _I64, _P, _U64, _P, _I32, _I32]),
    "sai_zt_isa": (C.c_char_p, []),
    "sai_zt_pack_i8": (_I64, [_LAY, _P, _P, _I64, _P, _U64, _P, _I32]),
    "sai_zt_decode_host": (C.c_int, [_LAY, _P, _P, _I64, _P]),
    "sai_zt_decode": (C.c_int, [_LAY, _P, _U64, _P, _I64, _I64, _P, _P]),
    "sai_synth_fill": (C.c_int, [_LAY, _P, _I64, _I64, _I64, _P, _U64, C.c_double, _P]),
}

_lib = None


def lib_path() -> Path:
    return _build.LIB


def load() -> C.CDLL:
    """Loads (building if necessary) the C-ABI library and sets prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    import os

    override = os.environ.get("SAI_B200_LIB")  # tools/ A/B scripts: the -DSAI_EXPERIMENTS build
    path = Path(override) if override else _build.build()
    lib = C.CDLL(str(path))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SaiError(RuntimeError):
    pass


def check(rc: int) -> None:
    """Maps a status code to the exception the reference raises for the same
    condition (``ValueError`` for argument errors)."""
    if rc == OK:
        return
    msg = load().sai_last_error().decode(errors="replace")
    if rc == E_ARG:
        raise ValueError(msg)
    if rc == E_DOMAIN:
        raise ValueError(msg)
    if rc == E_NOMEM:
        raise MemoryError(msg)
    raise SaiError(f"sai_b200 error {rc}: {msg}")
