"""Plain-Python mirrors of the reference's configuration objects.

Same accessors and validation behaviour as ``StatConfig``
(sai/configs/stat_config.py:36-226), ``PloidyConfig``
(sai/configs/ploidy_config.py:25-96) and the YAML loading of ``score``
(sai/sai.py:65-77), without the pydantic dependency.  A reference-side
integration passes its own pydantic objects instead: only ``.root``,
``.get_parameters()`` and ``.get_ploidy()`` are used by the GPU path.
"""

from __future__ import annotations

from typing import Any, Optional, Union

SUPPORTED_STATISTICS = ["Danc", "DD", "df", "Dplus", "fd", "U", "Q"]
_COMPARATORS = ["<=", ">=", "=", "<", ">"]


def parse_comparator(value: str, stat_name: str, param: str) -> tuple[str, float]:
    """``"=1"`` -> ``("=", 1.0)``; same lookup order and errors as
    ``StatConfig.check_comparator`` (stat_config.py:159-207)."""
    if not any(c in value for c in _COMPARATORS):
        raise ValueError(
            f"{param} for {stat_name} must contain a valid comparator (e.g., '=0.5', '>=0.2')."
        )
    comp = next(c for c in _COMPARATORS if c in value)
    try:
        num = float(value[len(comp) :])
    except ValueError:
        raise ValueError(
            f"{param} value for {stat_name} must be a valid number after the comparator."
        )
    if not (0 <= num <= 1):
        raise ValueError(f"{param} value must be between 0 and 1 for {stat_name}, but got {num}.")
    return comp, num


class StatConfig:
    def __init__(self, root: dict[str, Any]):
        for name, params in root.items():
            if name not in SUPPORTED_STATISTICS:
                raise ValueError(f"The {name} statistic is not supported.")
            if name in ("U", "Q"):
                self._check_u_q(name, params)
        self.root = root

    @staticmethod
    def _check_u_q(name: str, params: dict) -> None:
        if set(params.keys()) != {"ref", "tgt", "src"}:
            raise ValueError(
                f"{name} must have exactly the keys: {{'ref', 'tgt', 'src'}}, but got {set(params.keys())}."
            )
        for group in ("ref", "tgt"):
            for pop, value in params[group].items():
                if not (0 <= float(value) <= 1):
                    raise ValueError(
                        f"{group}[{pop}] value must be between 0 and 1 for {name}, got {value}."
                    )
        parsed = {}
        for pop, expr in params["src"].items():
            if isinstance(expr, tuple):  # already normalised
                parsed[pop] = expr
                continue
            if not isinstance(expr, str):
                raise ValueError(f"src[{pop}] value must be a comparator string for {name}.")
            parsed[pop] = parse_comparator(expr, name, f"src[{pop}]")
        params["src"] = parsed

    def get_parameters(self, stat_name: str):
        return self.root.get(stat_name, None)


class PloidyConfig:
    def __init__(self, root: dict[str, dict[str, int]]):
        allowed, required = {"ref", "tgt", "src", "outgroup"}, {"ref", "tgt", "src"}
        if set(root) - allowed:
            raise ValueError(
                f"Unsupported ploidy keys: {set(root) - allowed}. Allowed keys are {allowed}."
            )
        if required - set(root):
            raise ValueError(f"Missing required ploidy keys: {required - set(root)}.")
        for group, sub in root.items():
            if not isinstance(sub, dict):
                raise ValueError(f"Value for '{group}' must be a dictionary of population -> ploidy.")
            for pop, ploidy in sub.items():
                if not isinstance(ploidy, int) or isinstance(ploidy, bool) or ploidy <= 0:
                    raise ValueError(f"Ploidy for '{group}:{pop}' must be a positive integer.")
        self.root = root

    def get_ploidy(self, group: str, population: Optional[str] = None) -> Union[int, list[int], None]:
        if group not in self.root:
            if group == "outgroup":
                return None
            raise KeyError(f"Group '{group}' not found in configuration.")
        if population is None:
            return list(self.root[group].values())
        if population not in self.root[group]:
            raise KeyError(f"Population '{population}' not found under group '{group}'.")
        return self.root[group][population]


class PopConfig:
    def __init__(self, root: dict[str, str]):
        for key in ("ref", "tgt", "src"):
            if key not in root:
                raise ValueError(f"Missing required population key: {key}.")
        self.root = root

    def get_population(self, group: str) -> Optional[str]:
        return self.root.get(group, None)


class GlobalConfig:
    """``statistics`` / ``ploidies`` / ``populations`` sections of the YAML
    (sai/configs/global_config.py:29-99)."""

    def __init__(self, statistics: dict, ploidies: dict, populations: dict):
        self.statistics = StatConfig(statistics)
        self.ploidies = PloidyConfig(ploidies)
        self.populations = PopConfig(populations)


def load_config(path: str) -> GlobalConfig:
    """YAML -> GlobalConfig with the error behaviour of ``score``
    (sai/sai.py:65-73); duplicate keys are rejected like the reference's
    ``UniqueKeyLoader`` (sai/utils/unique_key_loader.py:26-72)."""
    import yaml

    class _Unique(yaml.SafeLoader):
        pass

    def _mapping(loader, node, deep=False):
        seen = set()
        for key_node, _ in node.value:
            key = loader.construct_object(key_node, deep=deep)
            if key in seen:
                raise ValueError(f"Duplicate key in YAML: {key!r}")
            seen.add(key)
        return yaml.SafeLoader.construct_mapping(loader, node, deep)

    _Unique.add_constructor(yaml.resolver.BaseResolver.DEFAULT_MAPPING_TAG, _mapping)
    try:
        with open(path, "r") as f:
            data = yaml.load(f, Loader=_Unique)
    except FileNotFoundError:
        raise FileNotFoundError(f"Configuration file '{path}' not found.")
    except yaml.YAMLError as e:
        raise ValueError(f"Error parsing YAML configuration file '{path}': {e}")
    for key in ("statistics", "ploidies", "populations"):
        if not isinstance(data, dict) or key not in data:
            raise ValueError(f"Configuration file '{path}' is missing the '{key}' section.")
    return GlobalConfig(data["statistics"], data["ploidies"], data["populations"])
