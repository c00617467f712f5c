#!/bin/bash
# ncu evidence for the kernels of the end-to-end call with the packers' zt wire (1 GPU): launch list of one bench run
# with a single e2e step (k_zt_decode + k_site per 32 MB slice), full capture of k_zt_decode.
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-strong --e2e-steps 1 --sites 2000000"
$CMD > $O/plain_zt.log 2>&1; echo "plain exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_zt_decode|k_site|k_window" -c 400 --csv --log-file $O/launches_zt_e2e.csv $CMD > $O/ncu_launches_zt.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_zt_decode -s 4 -c 2 -o $O/prof_k_zt_decode $CMD > $O/ncu_full_zt.log 2>&1
echo "k_zt_decode capture exit $?"
ls -la $O/*.ncu-rep | tail -3; tail -1 $O/plain_zt.log | cut -c1-300
