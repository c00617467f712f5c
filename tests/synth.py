"""Seeded synthetic genotype matrices for tests (numpy; CPU).

Model follows SURVEY.md 8(d): strictly increasing unique positions with
geometric gaps, per-site frequency ~ Beta(0.2, 2.0), a fraction of
"introgressed" sites (src fixed derived, ref ~ 0, tgt ~ U(0, 0.8)), optional
per-genotype missingness and sites missing in a whole source population.
Values are per-individual allele sums (0..ploidy), missing = -1 or -2.
"""

from __future__ import annotations

import numpy as np


def positions(rng, n_sites: int, mean_gap: float, start: int = 1) -> np.ndarray:
    gaps = rng.geometric(1.0 / mean_gap, size=n_sites).astype(np.int64)
    return (start - 1 + np.cumsum(gaps)).astype(np.int32)


def population(rng, freq: np.ndarray, n_ind: int, ploidy: int, missing: float = 0.0) -> np.ndarray:
    g = rng.binomial(ploidy, freq[:, None], size=(freq.shape[0], n_ind)).astype(np.int8)
    if missing > 0:
        m = rng.random(g.shape) < missing
        g[m] = np.where(rng.random(int(m.sum())) < 0.5, -1, -2).astype(np.int8)
    return g


def make_populations(
    seed: int,
    n_sites: int,
    pops: dict,
    mean_gap: float = 40.0,
    introgressed: float = 0.01,
    missing: float = 0.0,
    src_all_missing: float = 0.0,
    start: int = 1,
):
    """``pops = {"ref": {name: (n_ind, ploidy)}, "tgt": {...}, "src": {...}}``
    -> ``(pos, {"ref": {name: int8 matrix}, ...})``."""
    rng = np.random.default_rng(seed)
    pos = positions(rng, n_sites, mean_gap, start)
    f = rng.beta(0.2, 2.0, size=n_sites)
    intro = rng.random(n_sites) < introgressed
    f_ref = np.where(intro, 0.0005, f)
    f_tgt = np.where(intro, rng.random(n_sites) * 0.8, np.clip(f * (0.5 + rng.random(n_sites)), 0, 1))
    f_src = np.where(intro, 1.0, f)
    out = {"ref": {}, "tgt": {}, "src": {}}
    for name, (n, p) in pops["ref"].items():
        out["ref"][name] = population(rng, f_ref, n, p, missing)
    for name, (n, p) in pops["tgt"].items():
        out["tgt"][name] = population(rng, f_tgt, n, p, missing)
    for name, (n, p) in pops["src"].items():
        g = population(rng, f_src, n, p, missing)
        if src_all_missing > 0:
            g[rng.random(n_sites) < src_all_missing] = -2
        out["src"][name] = g
    if "outgroup" in pops:  # drawn last so that the other populations do not depend on it
        out["outgroup"] = {}
        f_out = np.where(rng.random(n_sites) < 0.03, 1.0, f * 0.05)
        for name, (n, p) in pops["outgroup"].items():
            out["outgroup"][name] = population(rng, f_out, n, p, missing)
    return pos, out
