"""sai_b200 -- B200-native U / Q95 sliding-window scoring for SAI.

Only what the hot path needs: the CUDA kernels and their C ABI (``csrc/``,
``include/sai_b200.h``), the host encoder, and mirrors of the reference's
``sai.stats`` / ``ChunkPreprocessor`` / ``score`` entry points.
"""

from .configs import GlobalConfig, PloidyConfig, PopConfig, StatConfig, load_config  # noqa: F401
from .encode import PackedGenotypes, PopData, make_layout, pack_populations, unpack_population  # noqa: F401
from .windows import chunk_windows, split_genome, split_windows_ranges  # noqa: F401

__version__ = "0.1.0"
