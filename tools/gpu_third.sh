#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== fps tests"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "fps" --timeout 120 > gpurun_out/pytest_fps.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_fps.log
echo "== bench"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['kernels_ms'])"
echo "== sweeps"
timeout 900 python bench.py --fps-sweep > gpurun_out/fps_sweep.log 2>&1; echo "fps sweep rc=$?"
