#!/bin/bash
# Second batch of round-1 evidence: bench with the zt wire, launch list including the e2e leg,
# full captures of the zt decoder and the DD kernels, host ingest throughput on the box's cores.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?"
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_e2e.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_zt_decode -s 40 -c 1 -o gpurun_out/prof_k_zt_decode $CMD > gpurun_out/ncu_full_zt.log 2>&1
echo "zt capture exit $?"
CMD6="python tests/run_configs.py --config 6"
$CMD6 > gpurun_out/config6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_site_hist|k_window_dd" -s 2 -c 2 -o gpurun_out/prof_dd $CMD6 > gpurun_out/ncu_full_dd.log 2>&1
echo "dd capture exit $?"
python tools/ingest_bench.py --sites 40000 > gpurun_out/ingest.log 2>&1; echo "ingest exit $?"
tail -1 gpurun_out/ingest.log
tail -1 gpurun_out/config6.log
tail -1 gpurun_out/bench.log | cut -c1-300
