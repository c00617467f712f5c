#!/bin/bash
# bench.py only, at 1/2/4/8 GPUs of one box (the driver's scaling run)
N=${1:-8}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_bench_n$n.log 2>&1
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_bench_n$n.log 2>&1
    fi
    echo "bench n=$n exit $?"
  fi
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_bench_n*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            e = d.get("e2e") or {}
            print(f, "value", round(d["value"] / 1e6, 2), "M/s step_ms", round(d["ms_per_step"], 4), "k1", round(d["roofline"]["k1_ms"], 4),
                  "e2e zt", round(e.get("value", 0) / 1e3, 1), "k/s dense", round((e.get("dense_tiles") or {}).get("value", 0) / 1e3, 1), "k/s ratio", round(e.get("wire_ratio", 0), 3))
PY
