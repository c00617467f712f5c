#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for rep in 1 2; do
for v in 0 1; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 0 --variant $v > gpurun_out/knob.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/knob.log"):
    if l.startswith("{"):
        d=json.loads(l); print("variant $v", round(d["value"]/1e6,2), "Mwin/s step", round(d["ms_per_step"],4), "k1", round(d["roofline"]["k1_ms"],4), "frac", round(d["roofline"]["frac"],4), "u", d["check"]["u_total"])
PY
done; done
