cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
for pf in 0 1; do
ALIFMM_DEBUG=1 timeout 600 python tests/probes/gpu_probe.py --nsrc 128 --check 0 --prefetch $pf 2>&1 | grep "source 0: rounds\|ttf wall\|sha1" | cut -c1-300
done
for pf in 0 1; do
ALIFMM_DEBUG=1 timeout 600 python tests/probes/gpu_probe.py --nsrc 16 --check 0 --prefetch $pf 2>&1 | grep "source 0: rounds\|ttf wall\|sha1" | cut -c1-300
done
