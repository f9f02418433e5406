set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_tests7.log
grep -E "passed|failed" gpurun_out/r2_tests7.log
export ALIFMM_DEBUG=1
for t in 768 1024 896 640; do
  echo "=== nsrc 128 threads $t"
  timeout 300 python tests/probes/gpu_probe.py --nsrc 128 --check 0 --reps 2 --threads $t --rays 8192 2>&1 | grep -E "cycles/round|slowest|ttf wall|rays wall" | tail -4
done > gpurun_out/r2_probe7.log 2>&1
cat gpurun_out/r2_probe7.log
unset ALIFMM_DEBUG
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_headline2.json 2> gpurun_out/bench_headline2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_headline2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['s_per_call'], {k:d['config'][k] for k in ('ms_seq_kernel','ms_march_kernel','ms_rays_kernel')}, d['roofline']['frac'], d['cpu_baseline']['value'])
"
