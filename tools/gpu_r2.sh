#!/bin/bash
# One gpurun call = several stages; every stage writes under gpurun_out/ and never aborts the rest.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2.sh newtests alltests bench sweep'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
for stage in "$@"; do
  echo "=== stage $stage ($(date +%T))"
  case "$stage" in
    newtests)
      timeout 600 python -m pytest tests/test_sa_fused_gpu.py tests/test_fp_and_concat_gpu.py tests/test_reference_callers_gpu.py tests/test_pointnet_util_gpu.py -q -m gpu 2>&1 | tail -40 | tee gpurun_out/newtests.log ;;
    alltests)
      timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 | tee gpurun_out/alltests.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 600 gpurun_out/bench_r2.err; head -c 1500 gpurun_out/bench_r2.json ;;
    benchquick)
      timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 400 gpurun_out/bench_quick.err
      python tools/show_bench.py gpurun_out/bench_quick.json ;;
    e2eprobe)
      for ex in 1 0; do for fl in 1 0; do PN2_SA_EXCLUSIVE=$ex PROBE_FLUSH=$fl timeout 300 python tools/e2e_probe.py; done; done 2>&1 | tee gpurun_out/e2e_probe.log | tail -40 ;;
    occupancy)
      timeout 300 python tools/cluster_occupancy.py 2>&1 | tee gpurun_out/cluster_occupancy.txt | tail -80 ;;
    fuzz)
      timeout 600 python tests/fuzz_gpu.py --seconds 150 --seed 7 --json gpurun_out/fuzz_r2.json 2>&1 | tail -4 ;;
    benchref)
      timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err; head -c 600 gpurun_out/bench_ref_r2.json ;;
    report)
      timeout 1200 python bench.py --report gpurun_out/report_r2.json > gpurun_out/report_r2.log 2>&1; tail -3 gpurun_out/report_r2.log ;;
    fpsbig)
      timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_parity_gpu.py tests/test_sa_fused_gpu.py -x -q -m gpu -k "fps or clouds" 2>&1 | tail -5
      timeout 120 python tools/cluster_occupancy.py > gpurun_out/fps_cluster_occupancy.txt 2>&1; tail -32 gpurun_out/fps_cluster_occupancy.txt
      timeout 600 python tools/fps_cluster_bench.py 2>&1 | tail -40
      timeout 60 tools/experiments/cluster_probe > gpurun_out/cluster_probe.txt 2>&1; grep -A9 "^launch" gpurun_out/cluster_probe.txt | head -70 ;;
    sweep)
      PN2_SWEEP_LARGE=1 timeout 900 python bench.py --fps-sweep > gpurun_out/fps_sweep_large.log 2>&1; cp gpurun_out/fps_sweep.json gpurun_out/fps_sweep_large_r2.json 2>/dev/null; tail -3 gpurun_out/fps_sweep_large.log ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | head -c 300 ;;
    ncufull)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fps_cta_kernel|ball_group_kernel' -s 4 -c 4 -o gpurun_out/r2_prof_layer python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | head -c 300 ;;
    prof_*)
      c=${stage#prof_}
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ball_group|fps_c|group_concat|group_rows|three_interp|fp_front|knn_kernel|inv_' -c 8 -f -o gpurun_out/r2_prof_$c python tools/prof_kernels.py $c 3 > gpurun_out/prof_$c.log 2>&1; tail -2 gpurun_out/prof_$c.log | head -c 300 ;;
    sanitize)
      timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/sanitize_mem.log 2>&1; tail -3 gpurun_out/sanitize_mem.log
      timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/sanitize_race.log 2>&1; tail -3 gpurun_out/sanitize_race.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
echo "=== done ($(date +%T))"
