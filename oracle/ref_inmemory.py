"""TEST / BENCH INFRASTRUCTURE -- module-level (hence picklable) subclasses of the reference's
`DataPreprocessor` / `DataGenerator` ABCs used by oracle/ref_driver.py.  Import only through
`ref_driver.make_classes()` (the reference package must be importable first)."""

from __future__ import annotations

from itertools import combinations
from typing import Any

import numpy as np

from ref_driver import _DATA

from sai.generators import DataGenerator
from sai.generators.chunk_generator import ChunkGenerator
from sai.generators.window_generator import WindowGenerator
from sai.preprocessors import DataPreprocessor
from sai.preprocessors.feature_preprocessor import FeaturePreprocessor
from sai.utils import split_genome
from sai.utils.genomic_dataclasses import ChromosomeData

class InMemoryChunkPreprocessor(DataPreprocessor):
    def __init__(self, key, win_len, win_step, ploidy_config, stat_config, anc_allele_available, output_file):
        self.key = key
        self.win_len, self.win_step = win_len, win_step
        self.ploidy_config = ploidy_config
        self.feature_preprocessor = FeaturePreprocessor(
            output_file=output_file, stat_config=stat_config, anc_allele_available=anc_allele_available)
        self.results = None

    def run(self, chr_name: str, start: int, end: int) -> list[dict[str, Any]]:
        d = _DATA[self.key]
        pos = d["pos"]
        lo, hi = np.searchsorted(pos, start, "left"), np.searchsorted(pos, end, "right")  # region chr:start-end

        def region(group):
            if not group:
                return None
            return {p: ChromosomeData(POS=pos[lo:hi], REF=None, ALT=None, GT=m[lo:hi]) for p, m in group.items()}

        wg = object.__new__(WindowGenerator)  # __init__ = read_data (allel) + the assignments below
        wg.win_len, wg.win_step, wg.chr_name = self.win_len, self.win_step, chr_name
        wg.num_src = len(d["src"])
        wg.ploidy_config = self.ploidy_config
        empty = hi <= lo  # read_geno_data returns None for a region without records (utils.py:140-141)
        wg.ref_data, wg.tgt_data, wg.src_data = (None, None, None) if empty else (region(d["ref"]), region(d["tgt"]), region(d["src"]))
        wg.out_data = None if empty else region(d["out"])
        names = lambda g: {p: [f"{p}_{i}" for i in range(m.shape[1])] for p, m in g.items()} if g else None
        wg.ref_samples, wg.tgt_samples, wg.src_samples = names(d["ref"]), names(d["tgt"]), names(d["src"])
        wg.out_samples = names(d["out"])
        wg.src_combinations = list(combinations(wg.src_samples.keys(), wg.num_src))
        wg.tgt_windows = {  # window_generator.py:132-144
            t: split_genome(
                pos=(wg.tgt_data[t].POS if (start is None) and (end is None) else [start, end - self.win_len + self.win_step]),
                window_size=self.win_len, step_size=self.win_step, start=start)
            for t in wg.tgt_samples
        }
        items = []
        for item in wg.get():  # chunk_preprocessor.py:142-147
            items.extend(self.feature_preprocessor.run(**item))
        return items

    def process_items(self, items) -> None:
        # mp_pool hands over the list of per-chunk lists (mp_pool.py:70-73); keep them
        self.results = items

class WindowRangeGenerator(DataGenerator):
    """What ChunkGenerator.get() yields (chunk_generator.py:84-98), for a given window list, without
    the pysam scan of ChunkGenerator.__init__."""

    def __init__(self, chr_name, windows, num_chunks):
        self.chr_name = chr_name
        self.chunks = ChunkGenerator._split_windows_ranges(None, windows, num_chunks)

    def get(self):
        for start, end in self.chunks:
            yield {"chr_name": self.chr_name, "start": start, "end": end}

    def __len__(self):
        return len(self.chunks)
