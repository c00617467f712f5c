#!/bin/bash
# Host packer A/B (experiments build): software-prefetch hint (0 T0, 1 T1, 2 T2, 3 NTA, 4 page heads only) x distance.
export SAI_B200_LIB=tools/bin/libsai_b200_exp.so
for cfg in "0 4" "1 4" "2 4" "3 4" "4 4" "4 1" "1 8" "1 16" "2 16" "0 4"; do
  set -- $cfg
  SAI_PACK_HINT=$1 SAI_PACK_AHEAD=$2 timeout 300 python tools/pack_bench.py --all-only --sites 1500000 --threads 1 4 16 | tee -a gpurun_out/pack_hint.jsonl
done
