set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 600 python tests/probes/gpu_probe.py --nsrc 128 --check 0 --rays 16256 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rays" 2>&1 | tail -3
