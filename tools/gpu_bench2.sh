#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
tail -3 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","n_gpus","ms_per_step","gpu_launches")}); print(l["e2e"]); print(l["device_batches_in_flight"])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 | tail -1 | cut -c1-300
