#!/usr/bin/env python
"""Wall-clock of the drop-in `score()` on a VCF file (what a `sai score` user waits for).

    python tools/score_bench.py [--sites 40000] [--samples 2504] [--all-stats]

Writes a synthetic bgzipped VCF (diploid, phased GT only; site frequencies ~ Beta(0.2, 2), 0.5 %
introgressed sites), a config with U + Q (or all seven statistics), then times
`sai_b200.score.score` (signature of sai.sai.score, sai/sai.py:33-151) end to end: chromosome
scan, native VCF parse of every population, the int8 pipeline (pack | H2D | kernels) + D2H, item
dicts, TSV / .log writing.  Prints one JSON line with the phase breakdown (a GPU is required).
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sai_b200 import preprocessors, score as score_mod, vcf  # noqa: E402


class Timer:
    def __init__(self):
        self.t = {}

    def wrap(self, name, fn):
        def inner(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                self.t[name] = self.t.get(name, 0.0) + time.perf_counter() - t0

        return inner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=40000)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--all-stats", action="store_true")
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    n_ref, n_tgt = int(a.samples * 0.6), int(a.samples * 0.4) - 4
    n_src = a.samples - n_ref - n_tgt
    tokens = np.array(["0|0", "0|1", "1|0", "1|1"])
    with tempfile.TemporaryDirectory() as tmp:
        head = "##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"i{k}" for k in range(a.samples)) + "\n"
        pos = np.cumsum(rng.geometric(1 / 40.0, size=a.sites))
        lines = []
        for p in pos:
            intro = rng.random() < 0.005
            fq = rng.beta(0.2, 2.0)
            f_ind = np.full(a.samples, fq)
            if intro:
                f_ind[:n_ref], f_ind[n_ref:n_ref + n_tgt], f_ind[n_ref + n_tgt:] = 0.0005, rng.random() * 0.8, 1.0
            g = (rng.random(a.samples) < f_ind).astype(np.int8) * 2 + (rng.random(a.samples) < f_ind)
            lines.append(f"1\t{p}\t.\tA\tT\t.\tPASS\t.\tGT\t" + "\t".join(tokens[g]))
        text = (head + "\n".join(lines) + "\n").encode()
        vcf_path = os.path.join(tmp, "s.vcf.gz")
        vcf.write_bgzf(vcf_path, text)
        anc = os.path.join(tmp, "anc.bed")
        with open(anc, "w") as f:
            f.writelines(f"1\t{p - 1}\t{p}\tA\n" for p in pos)
        pops = {}
        for g, (lo, hi, name) in {"ref": (0, n_ref, "REF"), "tgt": (n_ref, n_ref + n_tgt, "TGT"), "src": (n_ref + n_tgt, a.samples, "SRC")}.items():
            pops[g] = os.path.join(tmp, f"{g}.list")
            with open(pops[g], "w") as f:
                f.writelines(f"{name}\ti{k}\n" for k in range(lo, hi))
        stats = {"U": {"ref": {"REF": 0.01}, "tgt": {"TGT": 0.5}, "src": {"SRC": "=1"}},
                 "Q": {"ref": {"REF": 0.01}, "tgt": {"TGT": 0.95}, "src": {"SRC": "=1"}}}
        if a.all_stats:
            stats.update({"Danc": True, "Dplus": True, "df": True, "fd": True, "DD": True})
        cfg = os.path.join(tmp, "cfg.yaml")
        with open(cfg, "w") as f:
            yaml.safe_dump({"statistics": stats, "ploidies": {"ref": {"REF": 2}, "tgt": {"TGT": 2}, "src": {"SRC": 2}}, "populations": pops}, f)
        out = os.path.join(tmp, "scores.tsv")
        score_mod.score(vcf_path, "1", 50000, 10000, anc, out, cfg, 1)  # warm-up: CUDA context, page cache
        tm = Timer()
        score_mod.ChunkGenerator = tm.wrap("chromosome_scan", score_mod.ChunkGenerator)
        vcf.read_data = tm.wrap("vcf_parse", vcf.read_data)
        # int8 matrices -> host pack | copy | kernels (pipelined) -> host results
        preprocessors.HostEngine.score_matrices = tm.wrap("gpu_score_matrices", preprocessors.HostEngine.score_matrices)
        preprocessors.HostEngine.pattern_sums = tm.wrap("gpu_pattern_sums", preprocessors.HostEngine.pattern_sums)
        preprocessors.HostEngine.dd_sums = tm.wrap("gpu_dd_sums", preprocessors.HostEngine.dd_sums)
        preprocessors.write_items = tm.wrap("write_tsv_logs", preprocessors.write_items)
        t0 = time.perf_counter()
        score_mod.score(vcf_path, "1", 50000, 10000, anc, out, cfg, 1)
        total = time.perf_counter() - t0
        n_windows = sum(1 for _ in open(out)) - 1
        size = len(text)
    print(json.dumps({
        "sites": a.sites, "samples": a.samples, "windows": n_windows, "vcf_text_bytes": size, "statistics": list(stats),
        "host_threads": os.cpu_count(), "score_s": total, "phases_s": {k: round(v, 4) for k, v in tm.t.items()},
        "other_python_s": round(total - sum(tm.t.values()), 4), "windows_per_s": n_windows / total,
        "genotypes_per_s": a.sites * a.samples / total,
    }))


if __name__ == "__main__":
    main()
