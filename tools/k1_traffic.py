#!/usr/bin/env python
"""ncu CSV (`--metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_site --csv`) ->
k1_traffic.json stamped with the hash of the k_site sources, so that bench.py only quotes
`roofline.traffic` for the kernel it actually times (a stale capture is refused there).

    python tools/k1_traffic.py gpurun_out/k1_traffic.csv gpurun_out/k1_traffic.json [n_sites]
"""

import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(src, dst, n_sites=6_000_000):
    from bench import k1_source_sha

    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append(r)
    per_launch = {}
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows:
        name = r.get("Metric Name", "")
        if not name.startswith("dram__bytes_"):
            continue
        val = float(r["Metric Value"].replace(",", "")) * unit_scale.get(r.get("Metric Unit", "byte"), 1.0)
        per_launch.setdefault(r["ID"], {"kernel": r["Kernel Name"]})[name] = val
    launches = [v for v in per_launch.values() if "dram__bytes_read.sum" in v and "dram__bytes_write.sum" in v]
    if not launches:
        raise SystemExit(f"no dram__bytes rows in {src}")
    reads = [v["dram__bytes_read.sum"] for v in launches]
    writes = [v["dram__bytes_write.sum"] for v in launches]
    out = {
        "kernel": launches[0]["kernel"], "n_sites": int(n_sites), "k1_source_sha": k1_source_sha(),
        "dram_bytes_per_launch": sum(a + b for a, b in zip(reads, writes)) / len(launches),
        "dram_bytes_read": reads, "dram_bytes_write": writes,
        "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, {len(launches)} launches of the benched build "
                  f"(tools/gpu/gpu_traffic.sh)",
    }
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 6_000_000)
