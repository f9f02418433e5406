cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
ALIFMM_DEBUG=1 timeout 600 python tests/probes/gpu_probe.py --nsrc 128 --check 0 2>&1 | grep "seq\|ttf wall\|sha1" | cut -c1-330
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
