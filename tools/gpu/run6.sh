set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 300 tools/bench/math_bench > gpurun_out/math_bench2.log 2>&1; grep -E "branch-light|glibc sin\+cos|glibc atan" gpurun_out/math_bench2.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_tests6.log
grep -E "passed|failed" gpurun_out/r2_tests6.log
export ALIFMM_DEBUG=1
for n in 128 16; do
  echo "=== nsrc $n"
  timeout 300 python tests/probes/gpu_probe.py --nsrc $n --frac 0.3 --check 0 --reps 2 --rays 8192 2>&1 | grep -E "cycles/round|slowest|ttf wall|rays wall"
done > gpurun_out/r2_probe6.log 2>&1
cat gpurun_out/r2_probe6.log
