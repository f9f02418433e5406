cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
for kb in 0 112 128 160; do
echo "== band_smem_kb $kb"
ALIFMM_DEBUG=1 timeout 300 python tests/probes/gpu_probe.py --nsrc 128 --check 0 --smemkb $kb --reps 2 2>&1 | grep "source 0: rounds\|ttf wall" | tail -2 | cut -c1-250
done
