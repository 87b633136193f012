#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
   tools/train_ddp_demo.py --model cls_ssg --steps 16 --batch 32 --json gpurun_out/train_demo_cls_ssg_2gpu.json 2>&1 | tail -5
timeout 600 python tools/train_ddp_demo.py --model cls_ssg --steps 16 --batch 32 --json gpurun_out/train_demo_cls_ssg_1gpu_b32.json 2>&1 | tail -2
