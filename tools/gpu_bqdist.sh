#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -q -m gpu -k "ball or bq or query" -x 2>&1 | tail -3
python tools/bq_dist_bench.py 2>&1 | tail -20
