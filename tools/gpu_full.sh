#!/bin/bash
# Full pass: all GPU tests, bench (both arms), report, sweeps, ncu launch list + full captures of the three step kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 1200 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
echo "== report + bq sweep"
timeout 900 python bench.py --report gpurun_out/report.json > gpurun_out/report.log 2>&1; echo "report rc=$?"
timeout 600 python bench.py --bq-sweep > gpurun_out/bq_sweep.log 2>&1; echo "bq sweep rc=$?"
echo "== ncu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_cta_kernel -s 2 -c 1 -o gpurun_out/prof_fps -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fps.log 2>&1; echo "ncu fps rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bq_grid_query_kernel -s 2 -c 1 -o gpurun_out/prof_bq -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bq.log 2>&1; echo "ncu bq rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_narrow -s 2 -c 1 -o gpurun_out/prof_group -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_group.log 2>&1; echo "ncu group rc=$?"
