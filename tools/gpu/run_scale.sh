# usage: run_scale.sh N
N=$1
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r02_headline_${N}gpu_strong.json 2> gpurun_out/bench_${N}gpu.err
tail -c 2500 gpurun_out/bench_r02_headline_${N}gpu_strong.json; tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300
