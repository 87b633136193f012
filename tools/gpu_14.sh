#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/group_tune.txt
for md in 0 1; do for g in 4 8 16 32; do echo "== mode $md ctas/sm $g" >> gpurun_out/group_tune.txt; PN2_GROUP_MODE=$md PN2_GROUP_CTAS=$g timeout 120 python tools/run_group.py 2>/dev/null >> gpurun_out/group_tune.txt; done; done
cat gpurun_out/group_tune.txt | grep -v torch | paste - - - - | tail -14
timeout 600 python -m pytest tests -q -m gpu --timeout 120 -k "group" 2>&1 | tail -2
