#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_window -s 3 -c 1 -o gpurun_out/prof_k_window $CMD > gpurun_out/ncu_full_win.log 2>&1
echo "window capture exit $?"
