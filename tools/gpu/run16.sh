cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r02_final.log 2>&1
tail -5 gpurun_out/gpu_tests_r02_final.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_headline_1gpu.json 2> gpurun_out/bench_headline.err
cut -c1-1200 gpurun_out/bench_r02_headline_1gpu.json; tail -3 gpurun_out/bench_headline.err
