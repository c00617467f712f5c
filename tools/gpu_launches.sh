#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
grep -E "k_site|k_window" gpurun_out/launches.csv | tail -8 | cut -d, -f5,13-
