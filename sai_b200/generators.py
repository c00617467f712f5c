"""Mirror of the reference's chunk generator (host geometry; exact integers).

``ChunkGenerator(vcf_file, chr_name, step_size, window_size, num_chunks)`` has the
constructor, ``get()`` and ``len()`` of ``sai.generators.ChunkGenerator``
(sai/generators/chunk_generator.py:34-108): first / last POS of the chromosome ->
window grid (``split_genome``) -> ``num_chunks`` contiguous window ranges, each reported as
``{"chr_name", "start", "end"}`` with ``start`` of its first and ``end`` of its last window
(so neighbouring chunks overlap by ``window_size - step_size``, pinned by the reference's
tests/generators/test_chunk_generator.py:39).  The pysam record loop is replaced by one
parallel native scan of the VCF text (``sai_vcf_chrom_span``).
"""

from __future__ import annotations

from typing import Iterator

from .vcf import chromosome_span
from .windows import split_genome, split_windows_ranges


class ChunkGenerator:
    def __init__(self, vcf_file: str, chr_name: str, step_size: int, window_size: int, num_chunks: int):
        span = chromosome_span(vcf_file, chr_name)
        if span is None:
            raise ValueError(f"Chromosome {chr_name} not found in VCF.")
        first_pos, last_pos, _ = span
        windows = split_genome([first_pos, last_pos], window_size, step_size)
        self.chunks = split_windows_ranges(windows, num_chunks)
        self.num_chunks = len(self.chunks)
        self.chr_name = chr_name

    def get(self) -> Iterator[dict]:
        for start, end in self.chunks:
            yield {"chr_name": self.chr_name, "start": start, "end": end}

    def __len__(self) -> int:
        return self.num_chunks
