cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "four_and_eight" > gpurun_out/strip_tests8.log 2>&1
tail -4 gpurun_out/strip_tests8.log
timeout 900 python tools/bench/split_bench.py --n 16384 --gpus 8 --out gpurun_out/split_bench_r02_16384_8gpu.json 2> gpurun_out/split8.err | cut -c1-1200
tail -2 gpurun_out/split8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --config vor4096 --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r02_vor4096_8gpu.json 2> gpurun_out/vor8.err
tail -c 1500 gpurun_out/bench_r02_vor4096_8gpu.json | cut -c1-1500
