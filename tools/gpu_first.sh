#!/bin/bash
# First GPU pass: smoke, parity tests, golden vectors, bench, sweeps, ncu. Every stage under its own timeout.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
echo "== cluster tests first (short timeout)"
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "every_kernel_variant" --timeout 60 > gpurun_out/pytest_cluster.log 2>&1; echo "cluster rc=$?"
tail -5 gpurun_out/pytest_cluster.log
echo "== all gpu tests"
timeout 1200 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
echo "== golden"
timeout 300 python oracle/gen_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"
echo "== bench"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
echo "== sweeps"
timeout 900 python bench.py --fps-sweep > gpurun_out/fps_sweep.log 2>&1; echo "fps sweep rc=$?"
timeout 600 python bench.py --bq-sweep > gpurun_out/bq_sweep.log 2>&1; echo "bq sweep rc=$?"
timeout 900 python bench.py --report gpurun_out/report.json > gpurun_out/report.log 2>&1; echo "report rc=$?"
echo "== ncu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_cta_kernel -s 2 -c 1 -o gpurun_out/prof_fps -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fps.log 2>&1; echo "ncu fps rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ball_query_kernel -s 2 -c 1 -o gpurun_out/prof_bq -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bq.log 2>&1; echo "ncu bq rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_point -s 2 -c 1 -o gpurun_out/prof_group -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_group.log 2>&1; echo "ncu group rc=$?"
ls -la gpurun_out
