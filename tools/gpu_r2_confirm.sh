#!/bin/bash
# Confirmation call on the final tree (packed FPS updates are the built-in default): GPU suite, smoke, the full
# default bench line, one ncu --set full capture of the layer's two kernels and the per-launch list of bench.py.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== tests ($(date +%T))"
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/confirm_alltests.log
echo "=== smoke ($(date +%T))"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/confirm_smoke.log
echo "=== bench ($(date +%T))"
timeout 400 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python tools/show_bench.py gpurun_out/r2_bench_final.json 2>&1 | head -12
echo "=== ncu full ($(date +%T))"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'fps_cta_kernel|ball_group_kernel' -c 4 -f -o gpurun_out/r2_prof_layer_packed python tools/prof_kernels.py layer_cfg2 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | head -c 300
echo "=== ncu launches ($(date +%T))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_packed.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | head -c 300
echo "=== done ($(date +%T))"
