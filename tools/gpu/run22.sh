set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out /tmp/ncu
# 0. the other BASELINE configs + the reference arm, with the final kernels
for cfg in nb weld1 fmc64 vor4096 big16384; do
  timeout 900 python bench.py --config $cfg --steps 2 --warmup 3 --parity 2 > gpurun_out/bench_r02_$cfg.json 2> gpurun_out/bench_$cfg.err
  tail -c 300 gpurun_out/bench_r02_$cfg.json | cut -c1-300; tail -2 gpurun_out/bench_$cfg.err | cut -c1-200
done
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_reference_arm.err
tail -c 400 gpurun_out/bench_r02_reference_arm.json
# 1. launch list of the bench command (serialised, cold cache: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log | cut -c1-300
cap() {  # name kernel-regex command...
  name=$1; regex=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page details > gpurun_out/${name}_full_r02.txt 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > /tmp/ncu/${name}_raw.csv 2>/dev/null
  python - <<PY
import csv
rows = list(csv.reader(open("/tmp/ncu/${name}_raw.csv", errors="replace")))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r or "ID" in r]
if hdr:
    h = rows[hdr[0]]; v = rows[-1]
    keep = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__cluster_x",
            "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "smsp__inst_executed_per_warp.ratio")
    out = {k: v[i] for i, k in enumerate(h) if k in keep and i < len(v)}
    import json; json.dump(out, open("gpurun_out/${name}_raw_r02.json", "w"), indent=1)
    print("${name}", out.get("gpu__time_duration.sum"), out.get("dram__bytes_read.sum"), out.get("dram__bytes_write.sum"))
PY
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source cuda,sass > /tmp/ncu/${name}_source.csv 2>/dev/null
  python tools/ncu_hot_lines.py /tmp/ncu/${name}_source.csv 30 > gpurun_out/${name}_hot_lines_r02.txt 2>&1
  head -3 gpurun_out/${name}_hot_lines_r02.txt | cut -c1-200
  rm -f /tmp/ncu/$name.ncu-rep
}
cap march 'ali_march_kernel' python tests/probes/gpu_probe.py --nsrc 128 --check 0
cap rays 'ali_rays_kernel' python tests/probes/gpu_probe.py --nsrc 16 --check 0 --rays 8192
cap cluster 'ali_march_cluster' python tests/probes/gpu_probe.py --nsrc 16 --check 0
ls -la gpurun_out | head -60
