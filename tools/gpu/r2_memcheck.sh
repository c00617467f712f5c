#!/bin/bash
# compute-sanitizer memcheck (ONE tool, small cases) over the kernels that are new or rewritten in round 2:
# pieces / quantile select / flags-from-counts / N3 / N4 / wide planes / the int8 pipeline.
O=gpurun_out; mkdir -p $O
K="column_quantiles or config4_sharded or site_counts_vs_oracle or multiallelic or dd_needs or dd_random or stat_cases or four_pop_stat or kat or pipeline_golden or widens or int8_pipeline_matches_packed_engine[1- or int8_pipeline_matches_packed_engine[33- or device_scorer_matches"
timeout 300 python -m pytest tests -m gpu -x -q -k "$K" > $O/memcheck_plain.log 2>&1; echo "plain exit $?" >> $O/memcheck_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/memcheck.log python -m pytest tests -m gpu -x -q -k "$K" > $O/memcheck_pytest.log 2>&1
echo "memcheck exit $?"
tail -3 $O/memcheck_plain.log; tail -4 $O/memcheck_pytest.log; grep -c "Invalid\|out of bounds\|misaligned" $O/memcheck.log; tail -5 $O/memcheck.log
