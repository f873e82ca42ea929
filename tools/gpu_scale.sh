#!/bin/bash
# 1 -> 8 GPU scaling of bench.py on one 8-GPU box, launched as the driver does; args: TAG
set -u
mkdir -p gpurun_out
TAG=${1:-scale}
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_gpus.txt
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
done
python - <<PY
import json
for N in (1, 2, 4, 8):
    try:
        d = json.loads(open("gpurun_out/${TAG}_bench_n%d.json" % N).read().strip().splitlines()[-1])
        c4 = d.get("config4_strong_split") or {}
        print(N, "weak %.2f M (%.3f ms)" % (d["value"] / 1e6, d["ms_per_step"]), "e2e %.2f M" % (d["e2e"]["value"] / 1e6),
              "config4 strong %.2f M (%.3f ms)" % (c4.get("value", 0) / 1e6, c4.get("ms_per_step", 0)), "iters max", d.get("iters_max"))
    except Exception as e:
        print(N, "no bench", e)
PY
