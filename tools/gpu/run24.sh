cd ${GRAFT_REPO_ROOT:-/root/repo}
ALIFMM_DEBUG=1 timeout 300 python tests/probes/gpu_probe.py --nsrc 128 --check 0 --reps 1 2>&1 | grep "source 0: rounds\|ttf wall" | tail -2 | cut -c1-250
