cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
run() { echo "== nsrc $1 cluster $2 threads $3"; ALIFMM_DEBUG=1 timeout 300 python tests/probes/gpu_probe.py --nsrc $1 --check 0 --cluster $2 --cthreads $3 2>&1 | grep "source 0: rounds\|ttf wall" | cut -c1-200; }
run 16 4 768
run 16 5 512
run 16 6 512
run 16 6 768
run 16 7 512
run 16 8 512
run 32 4 768
run 32 3 768
run 64 2 768
run 64 2 512
