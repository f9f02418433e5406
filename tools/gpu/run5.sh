set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 300 tools/bench/math_bench > gpurun_out/math_bench.log 2>&1; cat gpurun_out/math_bench.log
export ALIFMM_DEBUG=1
for lib in "" "ali_fmm_and_ray_tracing_b200/libalifmm_noinline.so"; do
  echo "=== lib [$lib] nsrc 128"
  ALIFMM_LIB=$lib timeout 300 python tests/probes/gpu_probe.py --nsrc 128 --frac 0.3 --check 0 --reps 2 --rays 8192 2>&1 | grep -E "cycles/round|slowest|ttf wall|rays wall"
done > gpurun_out/r2_probe5.log 2>&1
echo "=== seq 128 threads" >> gpurun_out/r2_probe5.log
timeout 300 python tests/probes/gpu_probe.py --nsrc 128 --frac 0.3 --check 0 --reps 2 --seqthreads 128 2>&1 | grep -E "slowest|ttf wall" >> gpurun_out/r2_probe5.log
cat gpurun_out/r2_probe5.log
unset ALIFMM_DEBUG
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; tail -5 gpurun_out/sanitizer_memcheck_smoke.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke.log 2>&1; tail -5 gpurun_out/sanitizer_racecheck_smoke.log
free -g | head -2; nproc
timeout 1200 python bench.py --config big16384 --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/bench_big16384.json 2> gpurun_out/bench_big16384.err
tail -c 2500 gpurun_out/bench_big16384.json; tail -5 gpurun_out/bench_big16384.err
