"""Whole-genome scoring on one or several GPUs: the window ranges of
``ChunkGenerator._split_windows_ranges`` (sai/generators/chunk_generator.py:111-142) applied to the
flattened (chromosome, window) list, each GPU scoring its share with ONE genotype pass and ONE
window launch.

The reference scores one chromosome per ``sai score`` call and would spread its windows over
workers as contiguous ranges whose site regions overlap by ``win_len - win_step`` (the halo,
tests/generators/test_chunk_generator.py:39).  Across a genome the same rule gives every rank a
list of *pieces* -- (chromosome, contiguous window range) -- and a piece needs the sites of
``[first_window.start, last_window.end]`` only.  The pieces of a rank are laid side by side in one
packed matrix (each starting on a tile boundary) so that the genotype pass streams them in one
launch; the window kernel is told which site range belongs to which window
(``sai_window_stats_pieces``), because positions restart on every chromosome.

No data-path collective: windows are independent units.  The only exchange of a genome run is the
``sai outlier`` threshold (``sai_b200.outlier.device_thresholds``: one all-gather).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _cabi
from .scoring import DeviceScorer, WindowResults

TILE = _cabi.TILE_SITES


@dataclass
class Piece:
    chrom: int  # index into the caller's chromosome list
    win_lo: int  # windows [win_lo, win_hi) of that chromosome's window list
    win_hi: int


def shard_genome(windows_per_chrom: Sequence[Sequence[tuple[int, int]]], world_size: int) -> list[list[Piece]]:
    """Pieces of every rank: the flattened (chromosome, window) list cut into ``world_size``
    contiguous ranges, ``len // n`` windows each and the first ``len % n`` one more -- exactly
    ``_split_windows_ranges`` (chunk_generator.py:130-141) -- then cut at chromosome boundaries."""
    counts = [len(w) for w in windows_per_chrom]
    total = sum(counts)
    base, extra = divmod(total, world_size)
    starts = np.concatenate([[0], np.cumsum(counts)])
    out, at = [], 0
    for r in range(world_size):
        nxt = at + base + (1 if r < extra else 0)
        pieces = []
        for c, n in enumerate(counts):
            lo, hi = max(at, int(starts[c])), min(nxt, int(starts[c + 1]))
            if hi > lo:
                pieces.append(Piece(c, lo - int(starts[c]), hi - int(starts[c])))
        out.append(pieces)
        at = nxt
    return out


def piece_site_range(pos: np.ndarray, windows: Sequence[tuple[int, int]], piece: Piece, align: bool = True) -> tuple[int, int]:
    """Site index range ``[lo, hi)`` of the chromosome that a piece needs: the sites inside
    ``[first_window.start, last_window.end]`` (a region read ``chr:start-end``, utils.py:119-122);
    ``lo`` rounded down to a tile boundary when ``align`` (extra leading sites are harmless: a
    window only sees the positions inside its own bounds)."""
    first, last = windows[piece.win_lo], windows[piece.win_hi - 1]
    lo = int(np.searchsorted(pos, first[0], "left"))
    hi = int(np.searchsorted(pos, last[1], "right"))
    if align:
        lo = lo // TILE * TILE
    return lo, max(lo, hi)


class GenomeBatch:
    """The pieces of one rank side by side on the device.

    ``tile0[p]`` is the first tile of piece ``p`` in the concatenated matrix, ``n_sites[p]`` its
    number of sites.  Fill ``d_packed`` / ``d_pos`` piece by piece (``packed_view(p)``,
    ``pos_view(p)``), then ``score(jobs)``."""

    def __init__(self, layout, piece_sites: Sequence[int], piece_windows: Sequence[Sequence[tuple[int, int]]], n_jobs: int,
                 device=None, cap_u: Optional[int] = None, cap_q: Optional[int] = None):
        import torch

        self.torch = torch
        self.layout = layout
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_sites = [int(n) for n in piece_sites]
        tiles = [(n + TILE - 1) // TILE for n in self.n_sites]
        self.tile0 = [0]
        for t in tiles:
            self.tile0.append(self.tile0[-1] + t)
        self.n_tiles = self.tile0[-1]
        self.tile_bytes = int(layout.pairs_per_site) * TILE * 8
        self.win_count = [len(w) for w in piece_windows]
        W = sum(self.win_count)
        self.W = W
        ws = np.fromiter((w[0] for wl in piece_windows for w in wl), dtype=np.int64, count=W)
        we = np.fromiter((w[1] for wl in piece_windows for w in wl), dtype=np.int64, count=W)
        first = np.concatenate([np.full(n, self.tile0[p] * TILE, dtype=np.int32) for p, n in enumerate(self.win_count)] or [np.zeros(0, np.int32)])
        last = np.concatenate([np.full(n, self.tile0[p] * TILE + self.n_sites[p], dtype=np.int32) for p, n in enumerate(self.win_count)] or [np.zeros(0, np.int32)])
        d = self.device
        self.d_packed = torch.empty(max(1, self.n_tiles * self.tile_bytes), dtype=torch.uint8, device=d)
        # padding sites at the end of a piece's last tile are never searched (they lie beyond the
        # piece's site range); keep their positions at INT32_MAX anyway
        self.d_pos = torch.full((max(1, self.n_tiles * TILE),), 2**31 - 1, dtype=torch.int32, device=d)
        self.d_ws, self.d_we = torch.from_numpy(ws).to(d), torch.from_numpy(we).to(d)
        self.d_first, self.d_last = torch.from_numpy(first).to(d), torch.from_numpy(last).to(d)
        self.scorer = DeviceScorer(layout, self.n_tiles * TILE, W, n_jobs, device=d, cap_u=cap_u, cap_q=cap_q)

    def packed_view(self, p: int):
        return self.d_packed[self.tile0[p] * self.tile_bytes : self.tile0[p + 1] * self.tile_bytes]

    def pos_view(self, p: int):
        return self.d_pos[self.tile0[p] * TILE : self.tile0[p] * TILE + self.n_sites[p]]

    def load_piece(self, p: int, packed: np.ndarray, pos: np.ndarray) -> None:
        """Host-packed tiles (``pack_populations`` of the piece's sites) and positions -> device."""
        torch = self.torch
        view = self.packed_view(p)
        if packed.nbytes != view.numel():
            raise ValueError(f"piece {p}: {packed.nbytes} packed bytes, expected {view.numel()}")
        view.copy_(torch.from_numpy(np.ascontiguousarray(packed)))
        self.pos_view(p).copy_(torch.from_numpy(np.ascontiguousarray(pos, dtype=np.int32)))

    def score(self, jobs) -> None:
        """One genotype pass over all pieces + one window launch (stream-ordered, no host sync)."""
        sc = self.scorer
        if self.n_tiles == 0 or self.W == 0:
            return
        sc.site_flags(self.d_packed, jobs)
        sc.window_stats(self.d_pos, self.d_ws, self.d_we, jobs, self.d_first, self.d_last)

    def results(self) -> WindowResults:
        return self.scorer.results()

    def piece_slices(self) -> list[slice]:
        at, out = 0, []
        for n in self.win_count:
            out.append(slice(at, at + n))
            at += n
        return out


def synth_fill_piece(batch: GenomeBatch, p: int, chrom_site_lo: int, chrom_sites: int, roles: Sequence[int], seed: int,
                     missing_rate: float = 0.0) -> None:
    """Bench / test helper: fills piece ``p`` with the device generator's genotypes of the
    chromosome sites ``[chrom_site_lo, chrom_site_lo + n_sites[p])`` (tile-addressed: a site has
    the same genotypes whichever rank or piece holds it).  ``chrom_site_lo`` must be tile aligned."""
    import torch

    if chrom_site_lo % TILE:
        raise ValueError("chrom_site_lo must be a multiple of the tile size")
    lay = batch.layout
    t_chrom = chrom_site_lo // TILE
    n_tiles = batch.tile0[p + 1] - batch.tile0[p]
    if n_tiles == 0:
        return
    role = (C.c_int32 * len(roles))(*[int(r) for r in roles])
    base = batch.d_packed.data_ptr() + (batch.tile0[p] - t_chrom) * batch.tile_bytes  # where the chromosome's tile 0 would be
    _cabi.check(_cabi.load().sai_synth_fill(C.byref(lay), base, t_chrom, n_tiles, int(chrom_sites), role, int(seed),
                                            float(missing_rate), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
