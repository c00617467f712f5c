#!/bin/bash
# Quick check after a host-side change (1 GPU): int8 pipeline parity tests + bench without the CPU legs.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "int8 or i8 or pipeline or pack or host" > $O/pytest_quick.log 2>&1; echo "pytest exit $?" >> $O/pytest_quick.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-strong > $O/bench_quick.log 2> $O/bench_quick.err; echo "bench exit $?" >> $O/bench_quick.err
tail -3 $O/pytest_quick.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
e=d['e2e']; r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'k1', round(r['k1_ms'],4), 'frac', round(r['frac'],4))
print('e2e', round(e['value']), round(e['ms_per_step'],1), 'pack', round(e['pack_alone_ms'],1), 'pack_zt', round(e.get('pack_zt_alone_ms',0),1), e.get('pack_zt_reproduces_encode'), 'wire', round(e['wire_alone_ms'],1), 'ratio', round(e['pipeline_vs_slowest_stage'],3), 'zt', round(e['prepacked_zt']['value']), 'dense', round(e['prepacked_dense']['value']), e.get('matches_device_path'))
print('wire ratio', round(e.get('wire_ratio', 0), 2), 'h2d', e['h2d_bytes_per_step'], 'dense_wire', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in e.get('dense_wire', {}).items() if k != 'note'})
PY
tail -2 $O/bench_quick.err
