#!/bin/bash
# 1 GPU: all tests, default bench (pipelined e2e), bilinear line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_g1.json 2> gpurun_out/r02_bench_g1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_g1.err
timeout 600 python bench.py --steps 10 --warmup 3 --sample-mode bilinear --no-strong --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_g1_bilinear.json 2> gpurun_out/r02_bench_g1_bilinear.err; echo "bilinear rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_g1", "r02_bench_g1_bilinear"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()}, "roofline", round(d["roofline"]["frac"],3), d["roofline"].get("frac_k4_alone"))
        if d.get("strong"): print("  strong", round(d["strong"]["ms_per_step"],3))
        if d.get("e2e"): e=d["e2e"]; print("  e2e", round(e["ms_per_step"],2), round(e["ms_per_step_one_call_at_a_time"],2), round(e["ms_per_step_all_copied"],2), round(e["pcie_GBps_per_rank"],1))
        if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    except Exception as e: print(f, "ERR", e)
PY
