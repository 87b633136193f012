#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 ./tools/microbench/lat > gpurun_out/microbench_lat.txt 2>&1; cat gpurun_out/microbench_lat.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_bucket_kernel -s 2 -c 1 -o gpurun_out/prof_fps_bucket -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fps_bucket.log 2>&1; echo "ncu rc=$?"
