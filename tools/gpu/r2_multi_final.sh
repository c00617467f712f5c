#!/bin/bash
# Round 2, final multi-GPU call (gpurun --gpus 8): the NCCL outlier test + quantile / sharding tests, then bench.py at
# N = 8 / 4 / 2 with the final engine (headline + e2e + the config-4 strong-scaling record).
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests -m gpu -x -q -k "nccl or column_quantiles or config4 or int8" > $O/pytest_multi_final.log 2>&1; echo "pytest exit $?" >> $O/pytest_multi_final.log
for N in 8 4 2; do
  timeout 600 $TR --nproc-per-node $N --master-port $((29900+N)) bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 3 > $O/bench_final_n$N.log 2> $O/bench_final_n$N.err; echo "bench N=$N exit $?" >> $O/bench_final_n$N.err
done
tail -5 $O/pytest_multi_final.log
python - <<'PY'
import json
for n in (8, 4, 2):
    try:
        d = json.loads(open(f'gpurun_out/bench_final_n{n}.log').read().strip().splitlines()[-1])
        e = d['e2e'] or {}
        print(n, 'value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 3), 'traffic', d['roofline']['traffic'], 'e2e', round(e.get('value', 0)), 'e2e_ms', round(e.get('ms_per_step', 0), 1),
              'pack_ms', round(e.get('pack_alone_ms', 0), 1), 'thr', e.get('host_threads'), 'ratio', round(e.get('pipeline_vs_slowest_stage', 0), 3), 'zt', round(e.get('prepacked_zt', {}).get('value', 0)), 'dense', round(e.get('prepacked_dense', {}).get('value', 0)))
        print('  strong', json.dumps({k: d['strong'][k] for k in d['strong'] if k not in ('workload', 'threshold_how', 'per_rank_sites')}))
    except Exception as ex:
        print(n, 'parse failed', ex)
PY
tail -3 $O/bench_final_n8.err
