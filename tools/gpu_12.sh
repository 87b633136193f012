#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:three_nn_kernel -s 1 -c 1 -o gpurun_out/prof_three_nn -f python tools/run_three_nn.py > gpurun_out/ncu_three_nn.log 2>&1; echo "ncu rc=$?"
