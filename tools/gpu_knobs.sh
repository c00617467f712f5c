#!/bin/bash
mkdir -p gpurun_out
for h in 0 1 0 1; do
SAI_WIN_HINT=$h timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 0 > gpurun_out/knob.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/knob.log"):
    if l.startswith("{"):
        d=json.loads(l); print("hint $h", round(d["value"]/1e6,2), "Mwin/s step", round(d["ms_per_step"],4), "k1", round(d["roofline"]["k1_ms"],4), "win", round(d["ms_per_step"]-d["roofline"]["k1_ms"],4))
PY
done
