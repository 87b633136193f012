#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "fps" --timeout 120 -x > gpurun_out/pytest_fps.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_fps.log
PN2_SWEEP_LARGE=1 timeout 900 python bench.py --fps-sweep > gpurun_out/fps_sweep.log 2>&1; echo "fps sweep rc=$?"
