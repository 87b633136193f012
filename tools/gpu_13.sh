#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 -k "three or interp or fp_ or smoke" > gpurun_out/pytest_nn.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_nn.log
timeout 300 python tools/run_three_nn.py
