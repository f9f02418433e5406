set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -15 > gpurun_out/r2_tests8.log
grep -E "passed|failed|SKIP" gpurun_out/r2_tests8.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_headline3.json 2> gpurun_out/bench_headline3.err
tail -2 gpurun_out/bench_headline3.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
for f in ("bench_headline3", "bench_2gpu"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("n_gpus", "value", "ms_per_step", "scaling")}, "e2e", d["e2e"]["value"], d["e2e"]["s_per_call"],
              {k: d["config"].get(k) for k in ("ms_seq_kernel", "ms_march_kernel", "ms_rays_kernel", "march_ctas_per_source", "weak_value", "fields_rank0")})
    except Exception as e:
        print(f, "ERR", e)
PY
