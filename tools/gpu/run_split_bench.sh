# usage: run_split_bench.sh N   -- BASELINE config 5 through bench.py: one 16384^2 field as N row strips
N=$1
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --config big16384 --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_r02_big16384_${N}gpu.json 2> gpurun_out/bench_big_${N}gpu.err
tail -c 2200 gpurun_out/bench_r02_big16384_${N}gpu.json; tail -3 gpurun_out/bench_big_${N}gpu.err | cut -c1-300
