#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "cluster_capacity" --timeout 120 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:group_rows_vec4 -s 40 -c 1 -o gpurun_out/prof_group320 -f python tools/run_group.py > gpurun_out/ncu_group320.log 2>&1; echo "ncu group320 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:three_nn_kernel -s 3 -c 1 -o gpurun_out/prof_three_nn -f python tools/run_three_nn.py > gpurun_out/ncu_three_nn.log 2>&1; echo "ncu three_nn rc=$?"
tail -3 gpurun_out/ncu_group320.log
