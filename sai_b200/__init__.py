"""sai_b200 -- B200-native U / Q95 sliding-window scoring for SAI.

Only what the hot path needs: the CUDA kernels and their C ABI (``csrc/``,
``include/sai_b200.h``), the host ingest / encoder, and mirrors of the reference's entry points:

    sai_b200.score.score, sai_b200.outlier.outlier          sai/sai.py
    sai_b200.generators.ChunkGenerator                      sai/generators/chunk_generator.py
    sai_b200.preprocessors.ChunkPreprocessor                sai/preprocessors/chunk_preprocessor.py
    sai_b200.multiprocessing.mp_pool                        sai/multiprocessing/mp_pool.py
    sai_b200.stats.STAT_REGISTRY (U, Q, Danc, Dplus, df, fd, DD)   sai/stats/
    sai_b200.scoring.HostEngine / DeviceScorer, sai_b200.ops      the C ABI from Python / torch
"""

from .configs import GlobalConfig, PloidyConfig, PopConfig, StatConfig, load_config  # noqa: F401
from .encode import (  # noqa: F401
    PackedGenotypes, PopData, ZtGenotypes, compress, decompress, make_layout, pack_populations, unpack_population,
)
from .windows import chunk_windows, split_genome, split_windows_ranges  # noqa: F401

__version__ = "0.1.0"
