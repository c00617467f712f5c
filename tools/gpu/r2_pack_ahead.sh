#!/bin/bash
# Host packer A/B (experiments build): software-prefetch distance of the all-populations packer, in sites.
export SAI_B200_LIB=tools/bin/libsai_b200_exp.so
for a in 4 0 1 2 8 16 32 64 4; do
  SAI_PACK_AHEAD=$a timeout 300 python tools/pack_bench.py --all-only --sites 2000000 | tee -a gpurun_out/pack_ahead.jsonl
done
