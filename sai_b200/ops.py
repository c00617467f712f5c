"""``torch.ops.sai_b200.*``: the kernels of the hot path as PyTorch custom ops.

A thin registration over the C ABI (``include/sai_b200.h``) for callers that live in a
torch program: tensors own the memory, the ops run on torch's current stream, and the
``sai_layout`` / ``sai_job`` structs travel as CPU uint8 tensors (their raw bytes).  No
computation happens here; ``DeviceScorer`` (scoring.py) calls the same entry points
directly through ctypes.

    import sai_b200.ops                      # registers the ops
    lay = ops.struct_tensor(layout); jb = ops.jobs_tensor(jobs)
    torch.ops.sai_b200.site_flags(packed, lay, jb, n_sites, mask_u, mask_q, qval, 0)
    torch.ops.sai_b200.window_stats(pos, ws, we, jb, mask_u, mask_q, qval, nsnps, u, q, q_cnt,
                                    u_start, q_start, totals, u_cand, q_cand)
    torch.ops.sai_b200.window_stats_pieces(pos, ws, we, first_site, last_site, jb, ...)   # several chromosomes, one launch
    torch.ops.sai_b200.column_quantiles(cols, 0.99, out)                                  # `sai outlier` thresholds
"""

from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch
from torch import Tensor

from . import _cabi


def struct_tensor(obj) -> Tensor:
    """Raw bytes of a ctypes struct / array as a CPU uint8 tensor."""
    return torch.from_numpy(np.frombuffer(bytes(obj), dtype=np.uint8).copy())


def jobs_tensor(jobs: Sequence["_cabi.Job"]) -> Tensor:
    if not 1 <= len(jobs) <= _cabi.MAX_JOBS:
        raise ValueError(f"between 1 and {_cabi.MAX_JOBS} jobs per call")
    return struct_tensor((_cabi.Job * len(jobs))(*jobs))


def _layout_of(t: Tensor) -> "_cabi.Layout":
    if t.device.type != "cpu" or t.dtype != torch.uint8 or t.numel() != C.sizeof(_cabi.Layout):
        raise ValueError("layout must be a CPU uint8 tensor of sizeof(sai_layout) bytes")
    return _cabi.Layout.from_buffer_copy(t.numpy().tobytes())


def _jobs_of(t: Tensor):
    n, rem = divmod(t.numel(), C.sizeof(_cabi.Job))
    if t.device.type != "cpu" or t.dtype != torch.uint8 or rem or not 1 <= n <= _cabi.MAX_JOBS:
        raise ValueError("jobs must be a CPU uint8 tensor of 1..8 sai_job structs")
    return (_cabi.Job * n).from_buffer_copy(t.numpy().tobytes()), n


def _stream(t: Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


@torch.library.custom_op("sai_b200::site_flags", mutates_args=("mask_u", "mask_q", "qval"), device_types="cuda")
def site_flags(packed: Tensor, layout: Tensor, jobs: Tensor, n_sites: int, mask_u: Tensor, mask_q: Tensor,
               qval: Tensor, variant: int) -> None:
    """K1: genotype pass + fused site conditions (sai_site_flags)."""
    lay = _layout_of(layout)
    jarr, n_jobs = _jobs_of(jobs)
    n_tiles = (n_sites + _cabi.TILE_SITES - 1) // _cabi.TILE_SITES
    if mask_u.shape != (n_jobs, max(1, n_tiles)) or mask_q.shape != mask_u.shape or qval.shape[0] != n_jobs:
        raise ValueError("mask_u / mask_q must be [n_jobs, n_tiles] int32 and qval [n_jobs, >= 32 * n_tiles] float64")
    with torch.cuda.device(packed.device):
        _cabi.check(_cabi.load().sai_site_flags(
            C.byref(lay), packed.data_ptr(), 0, n_tiles, n_tiles, jarr, n_jobs, mask_u.data_ptr(), mask_q.data_ptr(),
            qval.data_ptr(), qval.shape[1], None, None, 0, variant, _stream(packed)))


@torch.library.custom_op(
    "sai_b200::window_stats",
    mutates_args=("nsnps", "u", "q", "q_cnt", "u_start", "q_start", "totals", "u_cand", "q_cand"), device_types="cuda")
def window_stats(pos: Tensor, win_start: Tensor, win_end: Tensor, jobs: Tensor, mask_u: Tensor, mask_q: Tensor,
                 qval: Tensor, nsnps: Tensor, u: Tensor, q: Tensor, q_cnt: Tensor, u_start: Tensor, q_start: Tensor,
                 totals: Tensor, u_cand: Tensor, q_cand: Tensor) -> None:
    """K4: per-window N(Variants), U, Q and candidate lists (sai_window_stats)."""
    jarr, n_jobs = _jobs_of(jobs)
    with torch.cuda.device(pos.device):
        _cabi.check(_cabi.load().sai_window_stats(
            pos.data_ptr(), pos.numel(), win_start.data_ptr(), win_end.data_ptr(), win_start.numel(), jarr, n_jobs,
            mask_u.data_ptr(), mask_q.data_ptr(), qval.data_ptr(), qval.shape[1], nsnps.data_ptr(), u.data_ptr(),
            q.data_ptr(), q_cnt.data_ptr(), u_start.data_ptr(), q_start.data_ptr(), totals.data_ptr(),
            u_cand.data_ptr(), u_cand.shape[1], q_cand.data_ptr(), q_cand.shape[1], _stream(pos)))


@torch.library.custom_op(
    "sai_b200::window_stats_pieces",
    mutates_args=("nsnps", "u", "q", "q_cnt", "u_start", "q_start", "totals", "u_cand", "q_cand"), device_types="cuda")
def window_stats_pieces(pos: Tensor, win_start: Tensor, win_end: Tensor, win_first_site: Tensor, win_last_site: Tensor,
                        jobs: Tensor, mask_u: Tensor, mask_q: Tensor, qval: Tensor, nsnps: Tensor, u: Tensor, q: Tensor,
                        q_cnt: Tensor, u_start: Tensor, q_start: Tensor, totals: Tensor, u_cand: Tensor,
                        q_cand: Tensor) -> None:
    """K4 over several chromosome pieces laid side by side: window i is searched only among the
    sites [win_first_site[i], win_last_site[i]) of its own piece (sai_window_stats_pieces)."""
    jarr, n_jobs = _jobs_of(jobs)
    if win_first_site.dtype != torch.int32 or win_last_site.dtype != torch.int32 or \
            win_first_site.numel() != win_start.numel() or win_last_site.numel() != win_start.numel():
        raise ValueError("win_first_site / win_last_site must be int32, one per window")
    with torch.cuda.device(pos.device):
        _cabi.check(_cabi.load().sai_window_stats_pieces(
            pos.data_ptr(), pos.numel(), win_start.data_ptr(), win_end.data_ptr(), win_first_site.data_ptr(),
            win_last_site.data_ptr(), win_start.numel(), jarr, n_jobs, mask_u.data_ptr(), mask_q.data_ptr(),
            qval.data_ptr(), qval.shape[1], nsnps.data_ptr(), u.data_ptr(), q.data_ptr(), q_cnt.data_ptr(),
            u_start.data_ptr(), q_start.data_ptr(), totals.data_ptr(), u_cand.data_ptr(), u_cand.shape[1],
            q_cand.data_ptr(), q_cand.shape[1], _stream(pos)))


@torch.library.custom_op("sai_b200::column_quantiles", mutates_args=("out",), device_types="cuda")
def column_quantiles(cols: Tensor, quantile: float, out: Tensor) -> None:
    """N1: exact linear quantile of every score column (sai_column_quantiles).  ``cols`` is float64
    ``[n_chunks, n_columns, length]`` (the output of one all_gather_into_tensor; NaN = no value) or
    ``[n_columns, length]``; ``out`` float64 ``[n_columns, 4]`` = threshold (NaN if undefined),
    number of values, minimum, maximum."""
    if cols.dtype != torch.float64 or out.dtype != torch.float64 or cols.dim() not in (2, 3) or not cols.is_contiguous():
        raise ValueError("cols must be a contiguous float64 tensor [chunks, columns, length] or [columns, length]")
    n_chunks = cols.shape[0] if cols.dim() == 3 else 1
    n_cols, length = int(cols.shape[-2]), int(cols.shape[-1])
    if tuple(out.shape) != (n_cols, 4) or not out.is_contiguous():
        raise ValueError("out must be a contiguous float64 tensor [columns, 4]")
    with torch.cuda.device(cols.device):
        _cabi.check(_cabi.load().sai_column_quantiles(cols.data_ptr(), n_cols, n_chunks, n_cols * length, length, length,
                                                      float(quantile), out.data_ptr(), _stream(cols)))
