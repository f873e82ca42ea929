#!/bin/bash
# N-GPU bench exactly as the driver launches it; args: TAG N
set -u
mkdir -p gpurun_out
TAG=${1:-m}; N=${2:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 1200 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","scaling","accepted_frac","gpu_launches")}, d["e2e"]["value"], d["config"]["global_batch"])
except Exception as e: print("no bench", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/${TAG}_ref_n$N.json 2> gpurun_out/${TAG}_ref_n$N.err; echo "ref N=$N rc=$?"; tail -c 300 gpurun_out/${TAG}_ref_n$N.json
