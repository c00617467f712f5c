"""Multi-GPU execution of the scoring path: one process per GPU, windows
sharded by contiguous ranges, no data-path collective.

Sharding follows ``ChunkGenerator._split_windows_ranges``
(sai/generators/chunk_generator.py:111-142): ``len // n`` windows per shard,
the first ``len % n`` shards one more; shard ``g`` covers the sites in
``[first_window.start, last_window.end]`` so neighbouring shards overlap by
``win_len - win_step`` base pairs (the halo, pinned by the reference's
tests/generators/test_chunk_generator.py:39).  Each rank scores its shard with
its own ``ChunkPreprocessor`` (``run(chr_name, start, end)``); results are
concatenated in shard order, which is what the reference's multi-process
drivers do (sai/multiprocessing/mp_pool.py:70-73, mp_manager.py:182).
"""

from __future__ import annotations

from typing import Any, Optional, Sequence

from .genome import shard_genome
from .windows import split_genome, split_windows_ranges


def shard_ranges(first_pos: int, last_pos: int, win_len: int, win_step: int, world_size: int) -> list[tuple[int, int]]:
    """``(start, end)`` of every shard (fewer than ``world_size`` when there
    are fewer windows than ranks)."""
    windows = split_genome([first_pos, last_pos], win_len, win_step)
    return split_windows_ranges(windows, world_size)


def my_shard(ranges: Sequence[tuple[int, int]], rank: int) -> Optional[tuple[int, int]]:
    return ranges[rank] if rank < len(ranges) else None


def run_sharded(preprocessor, chr_name: str, first_pos: int, last_pos: int, win_len: int, win_step: int,
                group=None) -> list[dict[str, Any]]:
    """Every rank runs ``preprocessor.run`` on its shard; rank 0 returns the
    items of all shards in genome order (other ranks return their own)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return preprocessor.run(chr_name, *shard_ranges(first_pos, last_pos, win_len, win_step, 1)[0])
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shard = my_shard(shard_ranges(first_pos, last_pos, win_len, win_step, world), rank)
    items = preprocessor.run(chr_name, shard[0], shard[1]) if shard is not None else []
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(items, gathered, dst=0, group=group)  # host-side result rows, not the data path
    if rank == 0:
        return [it for part in gathered for it in part]
    return items


def run_genome_sharded(preprocessors: dict, spans: dict, win_len: int, win_step: int, group=None) -> list[dict[str, Any]]:
    """Whole-genome version of ``run_sharded``: ``preprocessors[chr_name]`` is the
    ``ChunkPreprocessor`` of that chromosome's VCF (the reference scores one chromosome per
    ``sai score`` call, sai/parsers/score_parser.py:90-96), ``spans[chr_name] = (first POS, last
    POS)`` (``ChunkGenerator`` finds them).  The flattened (chromosome, window) list is cut into
    ``world_size`` contiguous ranges exactly like ``_split_windows_ranges`` (``genome.shard_genome``);
    a rank runs ``run(chr_name, first_window.start, last_window.end)`` for each of its pieces --
    a region read plus one engine call per piece -- and rank 0 returns all items in genome order
    (chromosomes in the order of ``spans``), the other ranks their own.  With the rows still on the
    ranks, ``sai_b200.outlier.outlier(score_file_of_this_rank, prefix, q, group)`` gives every rank
    the genome-wide thresholds with one small all-gather per column."""
    import torch.distributed as dist

    names = list(spans)
    windows = [split_genome([int(spans[c][0]), int(spans[c][1])], win_len, win_step) for c in names]
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    items: list[dict[str, Any]] = []
    for piece in shard_genome(windows, world)[rank]:
        w = windows[piece.chrom]
        items.extend(preprocessors[names[piece.chrom]].run(names[piece.chrom], w[piece.win_lo][0], w[piece.win_hi - 1][1]))
    if not distributed:
        return items
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(items, gathered, dst=0, group=group)  # host-side result rows, not the data path
    if rank == 0:
        return [it for part in gathered for it in part]
    return items
