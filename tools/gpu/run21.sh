cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
for st in 64; do
echo "== seq_threads $st"
ALIFMM_DEBUG=1 timeout 300 python tests/probes/gpu_probe.py --nsrc 128 --check 0 --seqthreads $st --reps 2 2>&1 | grep "seq\|ttf wall\|sha1" | cut -c1-330
done
