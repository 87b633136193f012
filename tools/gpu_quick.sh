#!/bin/bash
# Quick confidence pass: the whole GPU suite (-x, as the driver runs it) and one bench line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g serial %.4g inflight %s"%(l["value"],l["e2e"]["value"],l["e2e"]["serial"]["value"],l["device_batches_in_flight"].get("value")))
print(l["roofline"]["secondary"])
PY
