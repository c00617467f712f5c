#!/bin/bash
# Round 2, fourth call (1 GPU): parity suite on the entry-driven DD pass + scratch pool, config 6 timings, bench, ncu of the DD kernels.
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest4.log 2>&1; echo "pytest exit $?" >> $O/pytest4.log
timeout 300 python tests/run_configs.py --config 6 > $O/config6_d.log 2>&1; echo "c6 exit $?" >> $O/config6_d.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench4.log 2> $O/bench4.err; echo "bench exit $?" >> $O/bench4.err
CMD="python tests/run_configs.py --config 6"
$CMD > $O/c6_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_site_dd|k_entry_dd|k_window_dd|k_window_patterns|k_site_products" -s 12 -c 5 -o $O/prof_n3n4_b $CMD > $O/ncu_n3n4_b.log 2>&1
echo "ncu exit $?"
tail -6 $O/pytest4.log; tail -2 $O/config6_d.log; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench4.log').read().strip().splitlines()[-1])
    e=d['e2e']; print('value', d['value'], 'e2e', e['value'], e['ms_per_step'], 'pack', e['pack_alone_ms'], e['pack_alone_gbps_int8'], 'wire', e['wire_alone_ms'], 'ratio', e['pipeline_vs_slowest_stage'])
    print('zt', e['prepacked_zt']['value'], 'dense', e['prepacked_dense']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['gpu_vs_cpu_arm'])
except Exception as ex:
    print('bench parse failed', ex)
PY
tail -3 $O/bench4.err; tail -2 $O/ncu_n3n4_b.log
