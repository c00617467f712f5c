#!/bin/bash
mkdir -p gpurun_out
for v in 0 8; do
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --variant $v"
ncu --set full --clock-control none --import-source on -k regex:k_site -s 3 -c 1 -o gpurun_out/prof_k1_v$v $CMD > gpurun_out/ncu_k1_v$v.log 2>&1
echo "v$v exit $?"
done
