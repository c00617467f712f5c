#!/bin/bash
# First-contact GPU run: parity tests, smoke, bench (both K1 variants).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
timeout 300 python bench.py --steps 20 --warmup 3 --variant 2 --no-cpu > gpurun_out/bench_v2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_v2.log
tail -25 gpurun_out/pytest.log; tail -5 gpurun_out/smoke.log; tail -3 gpurun_out/bench.log; tail -3 gpurun_out/bench_v2.log
