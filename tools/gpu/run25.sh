cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gpu_strips" > gpurun_out/strip_tests.log 2>&1
tail -6 gpurun_out/strip_tests.log
timeout 900 python tools/bench/split_bench.py --n 16384 --gpus 4 --out gpurun_out/split_bench_r02_16384_4gpu.json 2> gpurun_out/split4.err | cut -c1-1200
tail -2 gpurun_out/split4.err
