"""Stat-class level mirror of the reference interface, backed by the GPU.

``STAT_REGISTRY.get("U")`` / ``.get("Q")`` return classes with the constructor
of ``GenericStatistic`` (sai/stats/generic_statistic.py:34-76) and the
``compute(**kwargs) -> {"name", "value", "cdd_pos"}`` contract of
``UStatistic`` (sai/stats/u_statistic.py:37-99) and ``QStatistic``
(sai/stats/q_statistic.py:37-104).  One call = one window = one trip to the
GPU, which is far too fine-grained for throughput (the batched entry point is
``sai_b200.preprocessors.ChunkPreprocessor``); this level exists so that the
reference's own unit tests can be run against the CUDA path unchanged.
"""

from __future__ import annotations

from typing import Any, Callable, Optional

import numpy as np

from .encode import pack_populations
from .scoring import HostEngine, make_job


class GenericRegistry:
    """Name -> class registry (sai/registries/generic_registry.py:36-89)."""

    def __init__(self):
        self._registry: dict[str, type] = {}

    def register(self, name: str) -> Callable[[type], type]:
        def deco(cls: type) -> type:
            if name in self._registry:
                raise ValueError(f"{name!r} is already registered.")
            self._registry[name] = cls
            return cls

        return deco

    def get(self, name: str) -> type:
        if name not in self._registry:
            raise KeyError(f"No component registered under name '{name}'")
        return self._registry[name]

    def list_registered(self) -> list[str]:
        return list(self._registry.keys())


STAT_REGISTRY = GenericRegistry()
_engine: Optional[HostEngine] = None


def _default_engine() -> HostEngine:
    global _engine
    if _engine is None:
        _engine = HostEngine(0)
    return _engine


class GenericStatistic:
    STAT_NAME = ""

    def __init__(
        self,
        ref_gts: np.ndarray,
        tgt_gts: np.ndarray,
        ref_ploidy: int,
        tgt_ploidy: int,
        src_gts_list: list[np.ndarray],
        src_ploidy_list: list[int],
        out_gts: Optional[np.ndarray] = None,
        out_ploidy: Optional[int] = None,
    ):
        self.ref_gts = ref_gts
        self.tgt_gts = tgt_gts
        self.src_gts_list = src_gts_list
        self.out_gts = out_gts
        self.ref_ploidy = ref_ploidy
        self.tgt_ploidy = tgt_ploidy
        self.src_ploidy_list = src_ploidy_list
        self.out_ploidy = out_ploidy

    def _run(self, u: Optional[dict], q: Optional[dict], y_list, anc: bool):
        if len(self.src_gts_list) != len(y_list):
            raise ValueError("The length of src_gts_list and y_list must match.")
        n_src = len(self.src_gts_list)
        ploidy = [self.ref_ploidy, self.tgt_ploidy] + list(self.src_ploidy_list)[:n_src]
        mats = [np.asarray(self.ref_gts), np.asarray(self.tgt_gts)] + [np.asarray(g) for g in self.src_gts_list]
        n = mats[0].shape[0]
        # one window over synthetic positions 0..n-1; candidates come back as row indices
        pg = pack_populations(mats, ploidy, np.arange(n, dtype=np.int32))
        job = make_job(0, 1, list(range(2, 2 + n_src)), anc, u, q)
        return _default_engine().score(pg, [(0, max(n - 1, 0))], [job])


def _require(kwargs: dict, keys: list[str]) -> None:
    missing = [k for k in keys if k not in kwargs]
    if missing:
        raise ValueError(f"Missing required argument(s): {', '.join(missing)}")


@STAT_REGISTRY.register("U")
class UStatistic(GenericStatistic):
    STAT_NAME = "U"

    def compute(self, **kwargs) -> dict[str, Any]:
        _require(kwargs, ["pos", "w", "x", "y_list", "anc_allele_available"])
        pos = np.asarray(kwargs["pos"])
        res = self._run(
            {"w": kwargs["w"], "x": kwargs["x"], "y_list": kwargs["y_list"]},
            None,
            kwargs["y_list"],
            kwargs["anc_allele_available"],
        )
        idx = res.u_positions(0, 0).astype(np.intp)
        return {"name": self.STAT_NAME, "value": int(res.u[0, 0]), "cdd_pos": pos[idx]}


@STAT_REGISTRY.register("Q")
class QStatistic(GenericStatistic):
    STAT_NAME = "Q"

    def compute(self, **kwargs) -> dict[str, Any]:
        _require(kwargs, ["pos", "w", "y_list", "anc_allele_available", "quantile"])
        pos = np.asarray(kwargs["pos"])
        res = self._run(
            None,
            {"w": kwargs["w"], "quantile": kwargs["quantile"], "y_list": kwargs["y_list"]},
            kwargs["y_list"],
            kwargs["anc_allele_available"],
        )
        value = res.q[0, 0]
        if np.isnan(value):
            return {"name": self.STAT_NAME, "value": np.nan, "cdd_pos": np.array([])}
        idx = res.q_positions(0, 0).astype(np.intp)
        return {"name": self.STAT_NAME, "value": np.float64(value), "cdd_pos": pos[idx]}


@STAT_REGISTRY.register("DD")
class DdStatistic(GenericStatistic):
    """``compute()`` -> ``{"name": "DD", "value": [one float per source population]}``
    (sai/stats/dd_statistic.py:40-77); the integer distance sums come from the GPU."""

    STAT_NAME = "DD"

    def compute(self, **kwargs) -> dict[str, Any]:
        from .scoring import dd_values

        n_src = len(self.src_gts_list)
        ploidy = [self.ref_ploidy, self.tgt_ploidy] + list(self.src_ploidy_list)[:n_src]
        mats = [np.asarray(self.ref_gts), np.asarray(self.tgt_gts)] + [np.asarray(g) for g in self.src_gts_list]
        n = mats[0].shape[0]
        pg = pack_populations(mats, ploidy, np.arange(n, dtype=np.int32), keep_negatives=True)
        eng = _default_engine()
        eng.score(pg, [(0, max(n - 1, 0))], [make_job(0, 1, list(range(2, 2 + n_src)), True)])
        ref_sum, tgt_sum = eng.dd_sums(pg, 0, 1, list(range(2, 2 + n_src)))
        vals = dd_values(ref_sum, tgt_sum, mats[0].shape[1], mats[1].shape[1], [m.shape[1] for m in mats[2:]])
        return {"name": self.STAT_NAME, "value": [v[0] for v in vals]}


class _FourPopStatistic(GenericStatistic):
    """Shared body of Danc / Dplus / df / fd: ``compute()`` -> ``{"name", "value": [one float per
    source population]}`` (danc_statistic.py:62-83, dplus_statistic.py:63-86, df_statistic.py:62-84,
    fd_statistic.py:63-89); the seven site-pattern sums come from the GPU (one window over all sites)."""

    def compute(self, **kwargs) -> dict[str, Any]:
        from .scoring import four_pop_values

        n_src = len(self.src_gts_list)
        ploidy = [self.ref_ploidy, self.tgt_ploidy] + list(self.src_ploidy_list)[:n_src]
        mats = [np.asarray(self.ref_gts), np.asarray(self.tgt_gts)] + [np.asarray(g) for g in self.src_gts_list]
        out_idx = -1
        if self.out_gts is not None:  # no outgroup: frequency 0 (stat_utils.py:212-213)
            out_idx = len(mats)
            mats.append(np.asarray(self.out_gts))
            ploidy.append(self.out_ploidy)
        n = mats[0].shape[0]
        pg = pack_populations(mats, ploidy, np.arange(n, dtype=np.int32))
        eng = _default_engine()
        eng.score(pg, [(0, max(n - 1, 0))], [make_job(0, 1, list(range(2, 2 + n_src)), True)])
        vals = four_pop_values(eng.pattern_sums(pg, 0, 1, out_idx, list(range(2, 2 + n_src))))
        return {"name": self.STAT_NAME, "value": [vals[self.STAT_NAME][k][0] for k in range(n_src)]}


@STAT_REGISTRY.register("Danc")
class DancStatistic(_FourPopStatistic):
    STAT_NAME = "Danc"


@STAT_REGISTRY.register("Dplus")
class DplusStatistic(_FourPopStatistic):
    STAT_NAME = "Dplus"


@STAT_REGISTRY.register("df")
class DfStatistic(_FourPopStatistic):
    STAT_NAME = "df"


@STAT_REGISTRY.register("fd")
class FdStatistic(_FourPopStatistic):
    STAT_NAME = "fd"
