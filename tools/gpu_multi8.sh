#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "8gpu rc=$?"
tail -2 gpurun_out/bench_8gpu.err
python -c "
import json
for l in open('gpurun_out/bench_8gpu.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('n_gpus',d['n_gpus'],'value %.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],'ms/step',d['ms_per_step'], d['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_ref_8gpu.json 2> gpurun_out/bench_ref_8gpu.err; echo "ref 8gpu rc=$?"
head -c 200 gpurun_out/bench_ref_8gpu.json
