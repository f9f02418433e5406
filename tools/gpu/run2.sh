set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_tests2.log
tail -5 gpurun_out/r2_tests2.log
export ALIFMM_DEBUG=1
for cfg in "16 8 512 256" "16 4 512 256" "16 0 0 32" "16 0 0 128" "128 0 0 256" "128 0 0 32" "64 0 0 256" "32 0 0 256"; do
  set -- $cfg
  echo "=== nsrc $1 cluster $2 threads $3 seq $4"
  timeout 300 python tests/probes/gpu_probe.py --nsrc $1 --frac 0.3 --cluster $2 --cthreads $3 --seqthreads $4 --check 0 --reps 2 2>&1 | grep -v "^create"
done > gpurun_out/r2_probe2.log 2>&1
grep -E "^===|clusters of|^cluster|ttf wall|cycles/round|slowest" gpurun_out/r2_probe2.log
