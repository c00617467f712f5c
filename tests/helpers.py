"""Shared helpers for the test-suite (fixture loading, item comparison)."""

from __future__ import annotations

import copy
import glob
import json
import os

import numpy as np

from sai_b200.configs import PloidyConfig, StatConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def SimpleStats(root: dict) -> StatConfig:
    return StatConfig(copy.deepcopy(root))


def SimplePloidy(root: dict) -> PloidyConfig:
    return PloidyConfig(copy.deepcopy(root))


def pipe_case_names() -> list[str]:
    return sorted(os.path.basename(p)[len("pipe_") : -len(".json")] for p in glob.glob(os.path.join(GOLDEN, "pipe_*.json")))


def vcf_case_names() -> list[str]:
    return sorted(os.path.basename(p)[len("vcf_") : -len(".json")] for p in glob.glob(os.path.join(GOLDEN, "vcf_*.json")))


def load_pipe_case(name: str):
    case = json.load(open(os.path.join(GOLDEN, f"pipe_{name}.json")))
    arrs = np.load(os.path.join(GOLDEN, f"pipe_{name}.npz"))
    data = {"ref": {}, "tgt": {}, "src": {}}
    if "outgroup" in case["ploidies"]:
        data["outgroup"] = {}
    for g in data:
        for p in case["ploidies"][g]:
            data[g][p] = arrs[f"{g}__{p}"]
    return case, arrs["pos"], data


def check_items(got: list[dict], exp: list[dict], q_tol: float = 0.0, four_tol: float = 0.0) -> None:
    """Item-by-item comparison with the reference's outputs: everything exact;
    Q bit-exact when ``q_tol == 0`` else within ``q_tol`` absolute."""
    assert len(got) == len(exp), (len(got), len(exp))
    for n, (g, e) in enumerate(zip(got, exp)):
        for k in ("chr_name", "start", "end", "ref_pop", "tgt_pop", "out_pop", "nsnps"):
            assert g[k] == e[k], (n, k, g[k], e[k])
        assert list(g["src_pop_list"]) == e["src_pop_list"], n
        for s in ("U", "Q"):
            if s not in e:
                assert s not in g, (n, s)
                continue
            if e[s] == "nan":
                assert np.isnan(g[s]), (n, s, g[s])
            elif s == "U":
                assert int(g[s]) == e[s], (n, g[s], e[s])
            else:
                want = float.fromhex(e[s])
                if q_tol == 0.0:
                    assert float(g[s]).hex() == e[s], (n, float(g[s]), want)
                else:
                    assert abs(float(g[s]) - want) <= q_tol, (n, float(g[s]), want)
            assert [int(p) for p in g["cdd_pos"][s]] == e[s + "_pos"], (n, s)
        for s in ("Danc", "Dplus", "df", "fd", "DD"):
            if s not in e:
                assert s not in g, (n, s)
                continue
            gv = g[s] if isinstance(g[s], list) else [g[s]]
            ev = e[s] if isinstance(e[s], list) else [e[s]]
            assert isinstance(g[s], list) == isinstance(e[s], list) and len(gv) == len(ev), (n, s)
            for a, b in zip(gv, ev):
                if b == "nan":
                    assert np.isnan(a), (n, s, a)
                elif four_tol == 0.0 or s == "DD":  # DD is exact: integer sums, the reference's own float steps
                    assert float(a).hex() == b, (n, s, float(a), float.fromhex(b))
                else:
                    want = float.fromhex(b)
                    assert abs(float(a) - want) <= four_tol * max(1.0, abs(want)), (n, s, float(a), want)


# ---- fakes for the process-pool tests (module level: the pool uses the spawn start method) ----
class SquareProcessor:
    """``run(x)`` -> ``[x * x, worker device]``; ``process_items`` keeps what it is given."""

    def __init__(self, path=None):
        self.path = path

    def run(self, x: int) -> list:
        from sai_b200 import multiprocessing as smp

        if x < 0:
            raise ValueError(f"cannot process {x}")
        return [(x * x, smp._worker_device)]

    def process_items(self, results: list) -> None:
        self.final_results = results
        if self.path:
            with open(self.path, "w") as f:
                f.write(repr(results))


class ListGenerator:
    def __init__(self, data: list):
        self.data = data

    def get(self):
        for x in self.data:
            yield {"x": x}
