"""Window grid and window-range sharding (host side, exact integer logic).

Mirrors ``split_genome`` (sai/utils/utils.py:558-612), the chunk worker's
re-derivation of its windows (sai/generators/window_generator.py:132-144) and
``ChunkGenerator._split_windows_ranges`` (sai/generators/chunk_generator.py:111-142).
"""

from __future__ import annotations

from typing import Optional, Sequence


def split_genome(
    pos: Sequence[int], window_size: int, step_size: int, start: Optional[int] = None
) -> list[tuple[int, int]]:
    """Sliding windows ``(start, end)`` (both inclusive) covering ``pos``.

    The first window starts at ``(pos[0] + step) // step * step - size + 1``
    clamped below by ``start`` (default 1); windows advance by ``step`` while
    their start is ``<= pos[-1]``.  Same ``ValueError`` conditions as the
    reference.
    """
    if step_size <= 0 or window_size <= 0:
        raise ValueError("`step_size` and `window_size` must be positive integers.")
    if step_size > window_size:
        raise ValueError("`step_size` cannot be greater than `window_size`.")
    if len(pos) == 0:
        raise ValueError("`pos` array must not be empty.")
    first_pos, last_pos = int(pos[0]), int(pos[-1])
    lower = 1 if start is None else int(start)
    win_start = max((first_pos + step_size) // step_size * step_size - window_size + 1, lower)
    n = 0 if win_start > last_pos else (last_pos - win_start) // step_size + 1
    return [(win_start + i * step_size, win_start + i * step_size + window_size - 1) for i in range(n)]


def chunk_windows(start: int, end: int, win_len: int, win_step: int) -> list[tuple[int, int]]:
    """Windows of a chunk ``(start, end)`` as a worker derives them."""
    return split_genome([start, end - win_len + win_step], win_len, win_step, start=start)


def split_windows_ranges(windows: list[tuple[int, int]], num_chunks: int) -> list[tuple[int, int]]:
    """Contiguous window ranges, ``len // n`` each with the first ``len % n``
    one longer; a range is ``(first.start, last.end)`` so neighbouring ranges
    overlap by ``win_len - win_step`` base pairs (the halo)."""
    base, extra = divmod(len(windows), num_chunks)
    out, at = [], 0
    for i in range(num_chunks):
        nxt = at + base + (1 if i < extra else 0)
        if nxt > at:
            out.append((windows[at][0], windows[nxt - 1][1]))
        at = nxt
    return out
