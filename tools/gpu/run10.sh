set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -8
# two-GPU row strips (bounded: the kernels give up after ~30 s without an answer from the peer)
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_gpu or two_devices" > gpurun_out/two_gpu_tests.log 2>&1
tail -25 gpurun_out/two_gpu_tests.log
# rays memoisation: the ray tests + the headline line (1 GPU)
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rays" > gpurun_out/ray_tests.log 2>&1
tail -5 gpurun_out/ray_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_headline_memo.json 2> gpurun_out/bench_headline_memo.err
tail -c 1500 gpurun_out/bench_headline_memo.json
