#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/sanitize_racecheck.log
for tool in initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"
  tail -3 gpurun_out/sanitize_$tool.log
done
