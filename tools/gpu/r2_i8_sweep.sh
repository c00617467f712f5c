#!/bin/bash
# int8 pipeline A/B on the GPU box: slice size x ring depth x non-temporal stores (experiments build only).
export SAI_B200_LIB=tools/bin/libsai_b200_exp.so
for cfg in "32 4 1" "8 4 1" "4 8 1" "4 8 0" "8 4 0" "2 8 0" "16 4 0" "32 4 0"; do
  set -- $cfg
  SAI_I8_SLICE_MB=$1 SAI_I8_RING=$2 SAI_PACK_NT=$3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('slice_mb $1 ring $2 nt $3', 'e2e_ms', round(e['ms_per_step'],1), 'pack_ms', round(e['pack_alone_ms'],1), 'ratio', round(e['pipeline_vs_slowest_stage'],3), e['matches_device_path'])
"
done
