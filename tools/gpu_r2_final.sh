#!/bin/bash
# Last GPU call of round 2: probe the packed FPS chain against the plain one, then run the whole GPU suite, the smoke
# test and a short bench with whichever chain the probe selected (PN2_FPS_PACKED), so that one call both decides
# and validates.  Everything lands under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== probe ($(date +%T))"
timeout 240 python tools/fps_packed_probe.py 2>&1 | tee gpurun_out/fps_packed_probe.log | tail -30
dec=$(grep -o "DECISION packed=[01]" gpurun_out/fps_packed_probe.log | tail -1 | grep -o "[01]$")
cdec=$(grep -o "DECISION_CLUSTER packed=[01]" gpurun_out/fps_packed_probe.log | tail -1 | grep -o "[01]$")
export PN2_FPS_PACKED=${dec:-0}
export PN2_FPS_PACKED_CLUSTER=${cdec:-0}
echo "PN2_FPS_PACKED=$PN2_FPS_PACKED PN2_FPS_PACKED_CLUSTER=$PN2_FPS_PACKED_CLUSTER" | tee gpurun_out/final_decision.txt
echo "=== tests ($(date +%T))"
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/final_alltests.log
echo "=== smoke ($(date +%T))"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
echo "=== bench ($(date +%T))"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python tools/show_bench.py gpurun_out/final_bench.json 2>&1 | head -30
echo "=== done ($(date +%T))"
