set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_tests4.log
grep -E "passed|failed" gpurun_out/r2_tests4.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err
tail -c 3000 gpurun_out/bench_headline.json; tail -5 gpurun_out/bench_headline.err
for cfg in nb weld1 fmc64 vor4096; do
  timeout 900 python bench.py --config $cfg --steps 2 --warmup 1 --parity 2 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  tail -c 1500 gpurun_out/bench_$cfg.json; tail -3 gpurun_out/bench_$cfg.err
done
timeout 1500 python tools/parity_survey.py > gpurun_out/r2_parity_survey.log 2>&1
tail -3 gpurun_out/r2_parity_survey.log | cut -c1-300
