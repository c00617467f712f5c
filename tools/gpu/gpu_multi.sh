#!/bin/bash
# N-GPU run: parity tests on one GPU, then the bench under torchrun, then the reference arm.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_n$N.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/bench_ref.log
tail -6 gpurun_out/pytest.log; grep -v "^$" gpurun_out/bench_n$N.log | tail -4 | cut -c1-1500; tail -3 gpurun_out/bench_ref.log | cut -c1-1200
