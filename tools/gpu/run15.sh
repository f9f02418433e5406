cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
ALIFMM_DEBUG=1 timeout 600 python tests/probes/gpu_probe.py --nsrc 128 --check 0 2>&1 | grep "source 0: rounds\|ttf wall\|sha1" | cut -c1-300
ALIFMM_DEBUG=1 timeout 600 python tests/probes/gpu_probe.py --nsrc 16 --check 0 2>&1 | grep "source 0: rounds\|ttf wall\|sha1" | cut -c1-300
