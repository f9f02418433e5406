cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
timeout 1500 python tools/parity_survey.py --out gpurun_out/parity_survey_r02_final.json > gpurun_out/parity_survey_final.log 2>&1
tail -8 gpurun_out/parity_survey_final.log | cut -c1-300
