#!/bin/bash
# Round 2, first GPU contact: box facts, parity tests, smoke, bench (both arms), configs 3/5, ncu launch list + traffic.
O=gpurun_out; mkdir -p $O
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv; nproc; free -g | head -2; lscpu | egrep "Model name|Socket|Thread|Core|NUMA|Flags" | cut -c1-400; } > $O/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref exit $?" >> $O/bench_ref.err
timeout 300 python tests/run_configs.py --config 5 > $O/config5.log 2>&1; echo "c5 exit $?" >> $O/config5.log
timeout 300 python tests/run_configs.py --config 3 > $O/config3.log 2>&1; echo "c3 exit $?" >> $O/config3.log
timeout 300 python tests/run_configs.py --config 6 > $O/config6.log 2>&1; echo "c6 exit $?" >> $O/config6.log
bash tools/gpu/gpu_traffic.sh > $O/traffic.log 2>&1
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 1"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list exit $?"
tail -30 $O/pytest.log; tail -3 $O/smoke.log; cat $O/bench.log | cut -c1-3000; tail -3 $O/bench.err; cat $O/bench_ref.log | cut -c1-1500; tail -2 $O/bench_ref.err; tail -2 $O/config5.log; tail -2 $O/config3.log | cut -c1-1500; tail -2 $O/config6.log; tail -3 $O/traffic.log; cat $O/box.txt
