set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_gpu" > gpurun_out/two_gpu_tests.log 2>&1
tail -5 gpurun_out/two_gpu_tests.log
timeout 900 python tools/bench/split_bench.py --n 4096 --out gpurun_out/split_bench_4096.json 2> gpurun_out/split_4096.err | cut -c1-1500
timeout 1200 python tools/bench/split_bench.py --n 16384 --out gpurun_out/split_bench_16384.json 2> gpurun_out/split_16384.err | cut -c1-1500
tail -3 gpurun_out/split_16384.err
