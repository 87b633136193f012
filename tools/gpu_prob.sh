#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/golden
python oracle/gen_golden.py --prob-only gpurun_out/golden 2>&1 | tail -5
cp gpurun_out/golden/prob_*.npz tests/golden/
python -m pytest tests -q -m gpu -k "prob" 2>&1 | tail -5
python -m pytest tests -q -m "not gpu" -k "prob" 2>&1 | tail -3
