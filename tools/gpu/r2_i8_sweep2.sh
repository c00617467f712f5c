#!/bin/bash
# int8 pipeline A/B (experiments build): can small slices with ordinary stores keep the tiles in the last-level cache
# for the copy engine (fewer DRAM round trips) now that the per-slice bubble is gone?
export SAI_B200_LIB=tools/bin/libsai_b200_exp.so
for cfg in "32 4 1" "4 8 0" "2 8 0" "8 4 0" "4 4 0" "4 8 1" "32 4 1"; do
  set -- $cfg
  SAI_I8_SLICE_MB=$1 SAI_I8_RING=$2 SAI_PACK_NT=$3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('slice_mb $1 ring $2 nt $3', 'e2e_ms', round(e['ms_per_step'],1), 'pack_ms', round(e['pack_alone_ms'],1), 'ratio', round(e['pipeline_vs_slowest_stage'],3), e['matches_device_path'])
"
done
