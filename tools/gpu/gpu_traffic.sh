#!/bin/bash
# DRAM bytes of the benched k_site build (roofline.traffic): ncu metrics pass -> gpurun_out/k1_traffic.json,
# stamped with the hash of the kernel sources; copy it to profiles/k1_traffic.json afterwards.
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 --no-strong"
$CMD > gpurun_out/traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_site -s 3 -c 3 --csv \
    --log-file gpurun_out/k1_traffic.csv $CMD > gpurun_out/traffic_ncu.log 2>&1
echo "ncu exit $?"
python tools/k1_traffic.py gpurun_out/k1_traffic.csv gpurun_out/k1_traffic.json
