#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus 8): box facts, per-rank H2D bandwidth with 1/2/4/8 ranks copying at once,
# the NCCL outlier test, bench.py at N = 8 / 4 / 2 (headline + e2e + the config-4 strong-scaling record).
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{ nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv; nproc; free -g | head -2; lscpu | egrep "Model name|Socket|Core|NUMA"; nvidia-smi topo -m; } > $O/box8.txt 2>&1
for N in 1 2 4 8; do
  timeout 120 $TR --nproc-per-node $N --master-port $((29700+N)) tools/h2d_probe.py > $O/h2d_n$N.json 2> $O/h2d_n$N.err
done
timeout 300 python -m pytest tests -m gpu -x -q -k "nccl or int8 or four_pop or pipeline_golden or all_statistics" > $O/pytest_multi.log 2>&1; echo "pytest exit $?" >> $O/pytest_multi.log
for N in 8 4 2; do
  timeout 600 $TR --nproc-per-node $N --master-port $((29800+N)) bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 3 > $O/bench_n$N.log 2> $O/bench_n$N.err; echo "bench N=$N exit $?" >> $O/bench_n$N.err
done
timeout 300 python tests/run_configs.py --config 6 > $O/config6_b.log 2>&1
cat $O/h2d_n*.json; tail -6 $O/pytest_multi.log
python - <<'PY'
import json
for n in (8, 4, 2):
    try:
        d = json.loads(open(f'gpurun_out/bench_n{n}.log').read().strip().splitlines()[-1])
        e = d['e2e'] or {}
        print(n, 'value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(e.get('value', 0)), 'e2e_ms', round(e.get('ms_per_step', 0), 1),
              'pack_ms', round(e.get('pack_alone_ms', 0), 1), 'zt', round(e.get('prepacked_zt', {}).get('value', 0)), 'dense', round(e.get('prepacked_dense', {}).get('value', 0)))
        print('  strong', json.dumps({k: d['strong'][k] for k in d['strong'] if k not in ('workload', 'threshold_how')}))
    except Exception as ex:
        print(n, 'parse failed', ex)
PY
tail -3 $O/bench_n8.err; tail -2 $O/config6_b.log; head -12 $O/box8.txt
