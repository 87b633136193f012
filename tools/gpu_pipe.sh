#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for f in 2 4 6; do
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --in-flight $f > gpurun_out/bench_if$f.json 2> gpurun_out/bench_if$f.err; echo "rc=$?"
python - <<PY
import json
l=json.load(open("gpurun_out/bench_if$f.json"))
print("value %.4g  step %.4f ms | e2e %.4g %.4f ms | serial %.4g %.4f ms match=%s"%(l["value"],l["ms_per_step"],l["e2e"]["value"],l["e2e"]["ms_per_step"],l["e2e"]["serial"]["value"],l["e2e"]["serial"]["ms_per_step"],l["e2e"]["outputs_match_device_leg"]))
print(l["device_batches_in_flight"], l["gpu_launches"], l["clocks"])
PY
done
