#!/bin/bash
# Multi-GPU box: the end-to-end leg with the packers' zt wire (host threads = all of a rank's cores); NS="8" (default), "4", "2".
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in ${NS:-8}; do
  timeout 500 $TR --nproc-per-node $N --master-port $((29900+N)) bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 3 --no-strong > $O/bench_zt_n$N.log 2> $O/bench_zt_n$N.err; echo "bench N=$N exit $?" >> $O/bench_zt_n$N.err
done
python - <<'PY'
import json
import os
for n in [int(x) for x in os.environ.get('NS', '8').split()]:
    try:
        d = json.loads(open(f'gpurun_out/bench_zt_n{n}.log').read().strip().splitlines()[-1])
        e = d['e2e'] or {}
        print(n, 'value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(e.get('value', 0)), 'e2e_ms', round(e.get('ms_per_step', 0), 1),
              'pack_ms', round(e.get('pack_alone_ms', 0), 1), 'thr', e.get('host_threads'), 'ratio', round(e.get('pipeline_vs_slowest_stage', 0), 3),
              'dense_wire', round(e.get('dense_wire', {}).get('value', 0)), 'zt', round(e.get('prepacked_zt', {}).get('value', 0)), 'dense', round(e.get('prepacked_dense', {}).get('value', 0)), e.get('matches_device_path'))
    except Exception as ex:
        print(n, 'parse failed', ex)
PY
tail -3 $O/bench_zt_n*.err
