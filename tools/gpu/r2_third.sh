#!/bin/bash
# Round 2, third call (1 GPU): full parity suite on the rewritten N3 / N4 kernels, configs 6 / 7, packer, bench, ncu captures.
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest3.log 2>&1; echo "pytest exit $?" >> $O/pytest3.log
timeout 300 python tests/run_configs.py --config 6 > $O/config6_c.log 2>&1; echo "c6 exit $?" >> $O/config6_c.log
timeout 300 python tests/run_configs.py --config 7 > $O/config7.log 2>&1; echo "c7 exit $?" >> $O/config7.log
timeout 300 python tools/pack_bench.py > $O/pack_bench2.json 2> $O/pack_bench2.err
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench3.log 2> $O/bench3.err; echo "bench exit $?" >> $O/bench3.err
CMD="python tests/run_configs.py --config 6"
$CMD > $O/c6_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_site_hist|k_window_patterns|k_site_products|k_site_dd|k_window_dd" -s 10 -c 5 -o $O/prof_n3n4 $CMD > $O/ncu_n3n4.log 2>&1
echo "ncu n3n4 exit $?"
tail -14 $O/pytest3.log; tail -2 $O/config6_c.log; tail -2 $O/config7.log | cut -c1-1500; cat $O/pack_bench2.json; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench3.log').read().strip().splitlines()[-1])
    e=d['e2e']; print('value', d['value'], 'e2e', e['value'], e['ms_per_step'], 'pack', e['pack_alone_ms'], e['pack_alone_gbps_int8'], 'ratio', e['pipeline_vs_slowest_stage'], 'traffic', d['roofline']['traffic'], d['roofline']['frac'])
    print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['gpu_vs_cpu_arm'])
except Exception as ex:
    print('bench parse failed', ex)
PY
tail -3 $O/bench3.err; tail -2 $O/ncu_n3n4.log
