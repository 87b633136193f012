#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 -k "group or sample_and or msg or sa_module or smoke" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --bq-sweep > gpurun_out/bq_sweep.log 2>&1; echo "bq sweep rc=$?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['kernels_ms'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ball_query_kernel -s 2 -c 1 -o gpurun_out/prof_bq -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bq.log 2>&1; echo "ncu bq rc=$?"
