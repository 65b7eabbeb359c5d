#!/bin/bash
# 1 GPU: every GPU test, the driver's default bench command, extra bench lines, and the ncu evidence for profiles/
mkdir -p gpurun_out
KREGEX='regex:align_stats|remap_median|build_pair|bbox_init|backproject_filter|grid_from|grid_store|clear_units|mark_|tile_count|tile_scan|zero_accum|unit_prefix|accumulate_|finalize_kernel|merge_|unmark'
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu.log
timeout 300 python scripts/gpu/kbench.py cfg2 7 > gpurun_out/r02_kbench.json 2>/dev/null; cat gpurun_out/r02_kbench.json
DDN_LIB_PATH=$PWD/depthdensifier_b200/libddn_b200_bulk.so timeout 300 python scripts/gpu/kbench.py cfg2 7 > gpurun_out/r02_kbench_bulk.json 2>/dev/null; cat gpurun_out/r02_kbench_bulk.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_g1.json 2> gpurun_out/r02_bench_g1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_g1.err
timeout 600 python bench.py --steps 10 --warmup 3 --sample-mode bilinear --no-strong --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_g1_bilinear.json 2> gpurun_out/r02_bench_g1_bilinear.err; echo "bilinear rc=$?"
timeout 600 python bench.py --impl reference --workload cfg1 --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_cfg1.json 2> gpurun_out/r02_bench_reference_cfg1.err; echo "ref cfg1 rc=$?"
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-strong > gpurun_out/r02_bench_g1_cfg1.json 2> gpurun_out/r02_bench_g1_cfg1.err; echo "cfg1 rc=$?"
# ncu: launch list of one step, then full captures of the three dominant kernels (first launch of each)
timeout 900 ncu -k "$KREGEX" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-e2e > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu list rc=$?"
for kn in backproject_filter_kernel accumulate_points_kernel remap_median_kernel; do
  timeout 900 ncu -k regex:$kn -c 1 --set full --import-source on --clock-control none -o gpurun_out/r02_${kn} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-e2e > gpurun_out/r02_ncu_${kn}.log 2>&1; echo "ncu $kn rc=$?"
  ncu -i gpurun_out/r02_${kn}.ncu-rep --page raw --csv > gpurun_out/r02_${kn}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_${kn}.ncu-rep --page source --csv > gpurun_out/r02_${kn}_source.csv 2>/dev/null
done
python scripts/summarise_launches.py gpurun_out/r02_launches_cfg2.csv build_pair_tables 2 | tail -16
python - <<'PY'
import json
for f in ("r02_bench_g1", "r02_bench_g1_bilinear", "r02_bench_g1_cfg1", "r02_bench_reference_cfg1"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), d.get("stages_ms") and {k:round(v,3) for k,v in d["stages_ms"].items()}, d.get("roofline") and (round(d["roofline"]["frac"],3), d["roofline"].get("frac_k4_alone")))
        if d.get("strong"): print("  strong", round(d["strong"]["ms_per_step"],3))
        if d.get("e2e") and "ms_per_step" in d["e2e"]: e=d["e2e"]; print("  e2e", round(e["ms_per_step"],2), e.get("ms_per_step_one_call_at_a_time"), e.get("ms_per_step_all_copied"), e.get("pcie_GBps_per_rank"))
        if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    except Exception as e: print(f, "ERR", e)
PY
