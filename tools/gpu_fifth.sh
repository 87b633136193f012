#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== fps tests"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "fps" --timeout 120 -x > gpurun_out/pytest_fps.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_fps.log
echo "== bench"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['kernels_ms'])"
echo "== sweeps"
PN2_SWEEP_SMALL=1 timeout 900 python bench.py --fps-sweep > gpurun_out/fps_sweep.log 2>&1; echo "fps sweep rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_bucket_kernel -s 2 -c 1 -o gpurun_out/prof_fps_bucket -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fps_bucket.log 2>&1; echo "ncu rc=$?"
