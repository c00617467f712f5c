#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 0 --variant $v > gpurun_out/knob_v$v.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/knob_v$v.log"):
    if l.startswith("{"):
        d=json.loads(l); print("variant $v", round(d["value"]/1e6,2), "Mwin/s step", round(d["ms_per_step"],4), "k1", round(d["roofline"]["k1_ms"],4), "win", round(d["ms_per_step"]-d["roofline"]["k1_ms"],4), "frac", round(d["roofline"]["frac"],4))
PY
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
