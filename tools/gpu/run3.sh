set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -120 > gpurun_out/r2_tests3.log
tail -25 gpurun_out/r2_tests3.log
export ALIFMM_DEBUG=1
for cfg in "128" "16" "64"; do
  echo "=== nsrc $cfg"
  timeout 300 python tests/probes/gpu_probe.py --nsrc $cfg --frac 0.3 --check 0 --reps 2 --rays 1024 2>&1 | grep -v "^create"
done > gpurun_out/r2_probe3.log 2>&1
grep -E "^===|^cluster|ttf wall|cycles/round|slowest|rays wall" gpurun_out/r2_probe3.log
unset ALIFMM_DEBUG
timeout 1500 python tools/parity_survey.py > gpurun_out/r2_parity_survey.log 2>&1
tail -5 gpurun_out/r2_parity_survey.log | cut -c1-400
