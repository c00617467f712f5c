#!/bin/bash
# Round 2, second call: int8 pipeline tests, host packer throughput, bench with the new e2e, ncu of the flags kernel.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "int8 or score_entry or pipeline_golden or vcf_fixture or mp_pool or source_comb" > $O/pytest2.log 2>&1; echo "pytest exit $?" >> $O/pytest2.log
timeout 300 python tools/pack_bench.py > $O/pack_bench.json 2> $O/pack_bench.err
timeout 120 python tools/h2d_probe.py > $O/h2d_n1.json 2> $O/h2d_n1.err
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench2.log 2> $O/bench2.err; echo "bench exit $?" >> $O/bench2.err
CMD="python tests/run_configs.py --config 5"
$CMD > $O/c5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_flags_from_counts -s 12 -c 2 -o $O/prof_k_flags $CMD > $O/ncu_flags.log 2>&1
echo "ncu flags exit $?"
tail -15 $O/pytest2.log; cat $O/pack_bench.json; tail -2 $O/pack_bench.err; cat $O/h2d_n1.json; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench2.log').read().strip().splitlines()[-1])
    print(json.dumps(d['e2e'], indent=1)); print('value', d['value'], 'cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None)
except Exception as e:
    print('bench parse failed', e)
PY
tail -5 $O/bench2.err; tail -3 $O/ncu_flags.log
