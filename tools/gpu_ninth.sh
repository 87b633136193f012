#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PN2_FPS_CFG=128,32,-3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_prune_kernel -s 2 -c 1 -o gpurun_out/prof_fps_prune -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fps_prune.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_cta_kernel -s 2 -c 1 -o gpurun_out/prof_fps -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fps.log 2>&1; echo "ncu rc=$?"
