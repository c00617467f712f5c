#!/bin/bash
# Host packer A/B on the box's CPUs (experiments build): row packer (4 = AVX-512BW mask tests, 5 = GFNI bit-matrix
# transpose), line writer (0 = 8-byte stores, 1 = 8x8 register transpose + 64-byte stores),
# prefetch distance.
export SAI_B200_LIB=tools/bin/libsai_b200_exp.so
grep -o -w "gfni\|avx512vbmi\|avx512_vbmi2\|avx512bw" /proc/cpuinfo | sort | uniq -c > gpurun_out/pack_ab_cpu.txt
for cfg in "4 0 4" "4 1 4" "5 0 4" "5 1 4" "5 1 2" "5 1 8" "5 1 16" "4 0 4" "5 1 4"; do
  set -- $cfg
  SAI_PACK_ISA=$1 SAI_PACK_FLUSH=$2 SAI_PACK_AHEAD=$3 timeout 300 python tools/pack_bench.py --all-only --sites 1500000 --threads 1 8 16 | tee -a gpurun_out/pack_ab.jsonl
done
