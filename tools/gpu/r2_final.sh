#!/bin/bash
# Round 2, evidence call (1 GPU): full parity suite, smoke, both bench arms, configs 3/5/6/7, host ingest / packer /
# score() timings, int8 pipeline slice A/B, ncu DRAM traffic of k_site and the launch list of a bench step.
O=gpurun_out; mkdir -p $O
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc; free -g | head -2; } > $O/box_final.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > $O/pytest_final.log 2>&1; echo "pytest exit $?" >> $O/pytest_final.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_final.log 2>&1; echo "smoke exit $?" >> $O/smoke_final.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_final.log 2> $O/bench_final.err; echo "bench exit $?" >> $O/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_final.log 2> $O/bench_ref_final.err; echo "ref exit $?" >> $O/bench_ref_final.err
for c in 3 5 6 7; do timeout 300 python tests/run_configs.py --config $c > $O/config${c}_final.log 2>&1; echo "c$c exit $?" >> $O/config${c}_final.log; done
timeout 300 python tools/pack_bench.py > $O/pack_bench_final.json 2> /dev/null
timeout 300 python tools/ingest_bench.py > $O/ingest_final.json 2> $O/ingest_final.err
timeout 300 python tools/score_bench.py > $O/score_bench_uq.json 2> $O/score_bench.err
timeout 300 python tools/score_bench.py --all-stats > $O/score_bench_all.json 2>> $O/score_bench.err
bash tools/gpu/gpu_traffic.sh > $O/traffic_final.log 2>&1
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --e2e-steps 1"
$CMD > $O/plain_final.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_final.csv $CMD > $O/ncu_launches_final.log 2>&1
echo "launch list exit $?"
( export SAI_B200_LIB=tools/bin/libsai_b200_exp.so
  for cfg in "32 4" "16 4" "8 6" "4 8" "64 3"; do
    set -- $cfg
    SAI_I8_SLICE_MB=$1 SAI_I8_RING=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('slice_mb $1 ring $2', 'e2e_ms', round(e['ms_per_step'],1), 'pack_ms', round(e['pack_alone_ms'],1), 'ratio', round(e['pipeline_vs_slowest_stage'],3), e['matches_device_path'])
"
  done ) > $O/i8_sweep_final.log 2>&1
tail -12 $O/pytest_final.log; tail -2 $O/smoke_final.log
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1])
    e=d['e2e']; r=d['roofline']
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'k1', round(r['k1_ms'],4), 'frac', round(r['frac'],4), 'traffic', r['traffic'], r['traffic_source'][:40])
    print('e2e', round(e['value']), round(e['ms_per_step'],1), 'pack', round(e['pack_alone_ms'],1), round(e['pack_alone_gbps_int8'],1), 'wire', round(e['wire_alone_ms'],1), 'ratio', round(e['pipeline_vs_slowest_stage'],3), 'zt', round(e['prepacked_zt']['value']), 'dense', round(e['prepacked_dense']['value']))
    print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['gpu_vs_cpu_arm'])
    print('strong', d['strong']['ms'], d['strong']['threshold_ms'], d['clocks'])
except Exception as ex:
    print('bench parse failed', ex)
PY
tail -2 $O/bench_final.err; cut -c1-600 $O/bench_ref_final.log; for c in 3 5 6 7; do tail -2 $O/config${c}_final.log | cut -c1-900; done
cat $O/pack_bench_final.json; cat $O/ingest_final.json; cat $O/score_bench_uq.json; cat $O/score_bench_all.json; tail -2 $O/score_bench.err; tail -2 $O/traffic_final.log | cut -c1-400; cat $O/i8_sweep_final.log
