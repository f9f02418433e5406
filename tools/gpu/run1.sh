set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_tests1.log
cat gpurun_out/r2_tests1.log | tail -15
export ALIFMM_DEBUG=1
for cfg in "16 1 768" "16 8 512" "16 8 768" "16 4 512" "32 4 512" "32 4 768" "32 2 768" "64 2 768" "64 2 512" "64 1 768"; do
  set -- $cfg
  echo "=== nsrc $1 cluster $2 threads $3"
  timeout 300 python tests/probes/gpu_probe.py --nsrc $1 --frac 0.3 --cluster $2 --cthreads $3 --check 0 --reps 2 2>&1 | grep -v "^create"
done > gpurun_out/r2_cluster_probe.log 2>&1
cat gpurun_out/r2_cluster_probe.log
