cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
for n in 16 64 128; do
echo "== nsrc $n seq_threads 32, reps 2"
ALIFMM_DEBUG=1 timeout 300 python tests/probes/gpu_probe.py --nsrc $n --check 0 --seqthreads 32 --reps 2 2>&1 | grep "slowest\|ttf wall" | cut -c1-200
done
