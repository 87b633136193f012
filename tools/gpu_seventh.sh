#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 1200 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
echo "== bench"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['kernels_ms'])"
echo "== report + bq sweep"
timeout 900 python bench.py --report gpurun_out/report.json > gpurun_out/report.log 2>&1; echo "report rc=$?"
timeout 600 python bench.py --bq-sweep > gpurun_out/bq_sweep.log 2>&1; echo "bq sweep rc=$?"
