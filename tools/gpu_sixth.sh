#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== fps tests"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "fps" --timeout 120 -x > gpurun_out/pytest_fps.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_fps.log
echo "== sweeps"
PN2_SWEEP_SMALL=1 timeout 900 python bench.py --fps-sweep > gpurun_out/fps_sweep.log 2>&1; echo "fps sweep rc=$?"
