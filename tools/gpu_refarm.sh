#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1]); print(l["value"], l["e2e"]["value"], l["cpu_baseline"])
r=json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1]); print(r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["single_core_value"])
PY
nproc; lscpu | grep -E "Model name|Thread|Core|Socket"
