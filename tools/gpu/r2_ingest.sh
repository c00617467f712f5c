#!/bin/bash
# Host ingest after the parser rewrite (box CPUs; score_bench needs the GPU): parser / packer / encoder throughput and
# the drop-in score() on a bgzipped VCF.
O=gpurun_out; mkdir -p $O
timeout 300 python tools/ingest_bench.py --sites 40000 > $O/ingest_r2b.json 2> $O/ingest_r2b.err; tail -1 $O/ingest_r2b.json | cut -c1-900
timeout 200 python tools/score_bench.py > $O/score_bench_uq3.json 2> $O/score_bench3.err; cat $O/score_bench_uq3.json | cut -c1-600
timeout 200 python tools/score_bench.py > $O/score_bench_uq3b.json 2>> $O/score_bench3.err; cat $O/score_bench_uq3b.json | cut -c1-600
