#!/bin/bash
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
timeout 900 python tests/run_configs.py --config 3 > gpurun_out/config3.log 2>&1; echo "exit $?" >> gpurun_out/config3.log
timeout 900 python tests/run_configs.py --config 5 > gpurun_out/config5.log 2>&1; echo "exit $?" >> gpurun_out/config5.log
timeout 900 python tests/run_configs.py --config 4 > gpurun_out/config4_n1.log 2>&1; echo "exit $?" >> gpurun_out/config4_n1.log
tail -4 gpurun_out/config3.log | cut -c1-1500; tail -4 gpurun_out/config5.log | cut -c1-1500; tail -4 gpurun_out/config4_n1.log | cut -c1-2500
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tests/run_configs.py --config 4 > gpurun_out/config4_n$N.log 2>&1; echo "exit $?" >> gpurun_out/config4_n$N.log
tail -4 gpurun_out/config4_n$N.log | cut -c1-2500
fi
