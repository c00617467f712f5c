#!/bin/bash
# int8 pipeline on the GPU box: parity of the host paths, then slice size x ring depth A/B (experiments build only).
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "int8 or pipeline_golden or vcf_fixture or score_entry or mp_pool or four_pop or source_comb or sharded" 2>&1 | tail -3
timeout 300 python tools/score_bench.py > $O/score_bench_uq.json 2> $O/score_bench.err; cat $O/score_bench_uq.json; tail -2 $O/score_bench.err
timeout 300 python tools/score_bench.py --all-stats > $O/score_bench_all.json 2>> $O/score_bench.err; cat $O/score_bench_all.json
export SAI_B200_LIB=tools/bin/libsai_b200_exp.so
for cfg in "32 4" "16 4" "8 6" "4 8" "64 3" "32 4"; do
  set -- $cfg
  SAI_I8_SLICE_MB=$1 SAI_I8_RING=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('slice_mb $1 ring $2', 'e2e_ms', round(e['ms_per_step'],1), 'pack_ms', round(e['pack_alone_ms'],1), 'ratio', round(e['pipeline_vs_slowest_stage'],3), e['matches_device_path'])
"
done
