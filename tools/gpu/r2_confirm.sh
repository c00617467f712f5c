#!/bin/bash
# Confirmation at HEAD (1 GPU): full parity suite, smoke, bench (both arms), score() timing.
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_confirm.log 2>&1; echo "pytest exit $?" >> $O/pytest_confirm.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_confirm.log 2>&1; echo "smoke exit $?" >> $O/smoke_confirm.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_confirm.log 2> $O/bench_confirm.err; echo "bench exit $?" >> $O/bench_confirm.err
timeout 300 python tools/score_bench.py > $O/score_bench_uq2.json 2> $O/score_bench2.err
timeout 300 python tools/score_bench.py --all-stats > $O/score_bench_all2.json 2>> $O/score_bench2.err
tail -4 $O/pytest_confirm.log; tail -2 $O/smoke_confirm.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_confirm.log').read().strip().splitlines()[-1])
e=d['e2e']; r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'k1', round(r['k1_ms'],4), 'frac', round(r['frac'],4), 'traffic', r['traffic'])
print('e2e', round(e['value']), round(e['ms_per_step'],1), 'pack', round(e['pack_alone_ms'],1), 'pack_zt', round(e.get('pack_zt_alone_ms',0),1), e.get('pack_zt_reproduces_encode'), 'wire', round(e['wire_alone_ms'],1), 'ratio', round(e['pipeline_vs_slowest_stage'],3), 'zt', round(e['prepacked_zt']['value']), 'dense', round(e['prepacked_dense']['value']))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['gpu_vs_cpu_arm'])
print('strong', d['strong']['ms'], d['strong']['threshold_ms'], d['clocks'], d['gpu_launches'])
PY
cat $O/score_bench_uq2.json $O/score_bench_all2.json; tail -2 $O/bench_confirm.err
