#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python tests/fuzz_gpu.py --seconds 120 --seed 2 --json gpurun_out/fuzz_seed2.json 2>&1 | tail -25
