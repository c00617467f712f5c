#!/bin/bash
# ncu evidence for the dominant kernel: launch list + one full capture.
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_site -s 3 -c 2 -o gpurun_out/prof_k_site $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_window -s 3 -c 1 -o gpurun_out/prof_k_window $CMD > gpurun_out/ncu_full_win.log 2>&1
echo "window capture exit $?"
tail -3 gpurun_out/plain.log; tail -5 gpurun_out/ncu_launches.log; tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out
