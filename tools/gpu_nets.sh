#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_nets_gpu.py -q -m gpu -x 2>&1 | tail -15
for mdl in cls_ssg sem_seg; do
  timeout 600 python tools/train_ddp_demo.py --model $mdl --steps 16 --batch 16 --json gpurun_out/train_demo_${mdl}_1gpu.json 2>&1 | tail -4
done
