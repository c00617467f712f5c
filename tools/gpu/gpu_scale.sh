#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_n$N.txt 2>&1
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_bench_n$n.log 2>&1
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_bench_n$n.log 2>&1
    fi
    echo "bench n=$n exit $?"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tests/run_configs.py --config 4 > gpurun_out/scale_config4_n$n.log 2>&1
    echo "config4 n=$n exit $?"
  fi
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_*_n*.log")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l)
            if "metric" in d: print(f, "value", round(d["value"]/1e6,2), "M/s step_ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]/1e3,1) if d.get("e2e") else None, "k/s")
            else: print(f, "windows/s", round(d["windows_per_s"]/1e6,2), "M max_rank_ms", round(d["max_rank_ms"],3), "thrU", d["threshold_u_q99"], "thrQ", d["threshold_q_q99"], "allgather_s", round(d["threshold_allgather_s"],4), "outliers", d["outliers_u"], d["outliers_q"], "u_sum", d["u_sum"])
PY
