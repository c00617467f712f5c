#!/bin/bash
# Round 2, ncu evidence of the final build (1 GPU): new custom-op test, clean launch list of a bench step (no e2e leg),
# full captures of k_site / k_window_stats and of the quantile select.
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "torch_ops or torch_custom or column_quantiles" 2>&1 | tail -3
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --e2e-steps 0"
$CMD > $O/plain_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_clean.csv $CMD > $O/ncu_launches_clean.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_site -s 3 -c 2 -o $O/prof_k_site_r2 $CMD > $O/ncu_full_ksite.log 2>&1
echo "k_site capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_window_stats -s 3 -c 1 -o $O/prof_k_window_r2 $CMD > $O/ncu_full_kwin.log 2>&1
echo "k_window capture exit $?"
CMD2="python tests/run_configs.py --config 4 --sites 20000000"
$CMD2 > $O/c4_plain.log 2>&1 && ncu --set full --clock-control none -k regex:k_column_quantile -s 1 -c 2 -o $O/prof_k_quantile $CMD2 > $O/ncu_quantile.log 2>&1
echo "quantile capture exit $?"
tail -2 $O/c4_plain.log | cut -c1-600; ls -la $O/*.ncu-rep
