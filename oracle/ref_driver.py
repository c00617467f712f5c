"""TEST / BENCH INFRASTRUCTURE -- drives the UNMODIFIED reference (oracle/_ref, installed by
oracle/build_ref.py) over in-memory genotype matrices.

Only tests/, smoke() and bench.py's CPU legs may import this module; the product path never does.

What runs is the reference's own code, unmodified:
    sai.multiprocessing.mp_pool.mp_pool                    mp_pool.py:45-73   (fork Pool.map)
    ChunkGenerator._split_windows_ranges                   chunk_generator.py:111-142
    WindowGenerator._window_generator / .get               window_generator.py:150-247, 290-307
    FeaturePreprocessor.run / STAT_REGISTRY / UStatistic / QStatistic / stat_utils
The one thing replaced is the VCF read: `WindowGenerator.__init__` calls `read_data` ->
`allel.read_vcf` (scikit-allel is not installed and VCF parsing is outside the timed path on both
arms), so `InMemoryChunkPreprocessor.run(chr_name, start, end)` -- a `DataPreprocessor` subclass
shaped like `ChunkPreprocessor.run` (chunk_preprocessor.py:105-147) -- fills the attributes
`__init__` would have set from a region read `chr:start-end` (window_generator.py:102-148) by
slicing the in-memory int64 matrices (the dtype the reference holds: `np.sum` of int8 alleles,
utils.py:405-410), then loops `for item in window_generator.get(): items.extend(fp.run(**item))`
exactly as chunk_preprocessor.py:142-147.
"""

from __future__ import annotations

import os
import sys
import types
from typing import Any, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
REF_DIR = os.path.join(HERE, "_ref")

_sai = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "sai", "stats", "stat_utils.py"))


def load():
    """Imports the reference package from oracle/_ref behind stubs for the three third-party
    modules that are not installed (never executed on this path)."""
    global _sai
    if _sai is not None:
        return _sai
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference is mounted")
    if "allel" not in sys.modules:
        allel = types.ModuleType("allel")
        allel.GenotypeVector = allel.GenotypeArray = object
        sys.modules["allel"] = allel
    sys.modules.setdefault("pysam", types.ModuleType("pysam"))
    if "natsort" not in sys.modules:
        ns = types.ModuleType("natsort")
        ns.natsorted = sorted
        sys.modules["natsort"] = ns
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import sai.stats  # noqa: F401  registers the statistics (sai/__main__.py:22)
    import sai

    _sai = sai
    return sai


# in-memory chromosome, set in the parent BEFORE mp_pool forks its workers; the pickled
# processor only carries the key
_DATA: dict[str, dict] = {}


def set_data(key: str, pos: np.ndarray, ref: dict, tgt: dict, src: dict, out: Optional[dict] = None) -> None:
    """``ref/tgt/src/out`` = {population: int64 matrix (sites x individuals)} over the same ``pos``."""
    _DATA[key] = dict(pos=np.asarray(pos), ref=ref, tgt=tgt, src=src, out=out)


def make_classes():
    """The in-memory processor / generator classes (module level in oracle/ref_inmemory.py so that
    mp_pool can pickle them) and the reference's split_genome."""
    load()
    import ref_inmemory as m

    return m.InMemoryChunkPreprocessor, m.WindowRangeGenerator, m.split_genome


def reference_configs(ploidies: dict, stats: dict):
    load()
    from sai.configs import PloidyConfig, StatConfig

    return PloidyConfig(ploidies), StatConfig(stats)


def score_windows(key: str, chr_name: str, windows: list, num_chunks: int, nprocess: int, win_len: int, win_step: int,
                  ploidies: dict, stats: dict, anc_allele_available: bool) -> list[dict[str, Any]]:
    """Scores ``windows`` (a contiguous run of the chromosome's window list) with the reference's
    `mp_pool` over `num_chunks` window-range chunks; returns the items in genome order.  With
    ``nprocess <= 1`` the chunks run serially in this process (what `score()` does, sai.py:148-149)."""
    Pre, Gen, _ = make_classes()
    from sai.multiprocessing.mp_pool import mp_pool

    pc, sc = reference_configs(ploidies, stats)
    pre = Pre(key, win_len, win_step, pc, sc, anc_allele_available, os.devnull)
    gen = Gen(chr_name, windows, max(1, min(num_chunks, len(windows))))
    if nprocess <= 1:
        parts = [pre.run(**params) for params in gen.get()]
    else:
        mp_pool(pre, gen, nprocess)
        parts = pre.results
    return [it for part in parts for it in part]
