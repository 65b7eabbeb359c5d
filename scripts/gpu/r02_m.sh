#!/bin/bash
# 1 GPU: parity after the K1K2 change + 1-GPU bench with the strong block (the N=1 reference of the strong scaling)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02m.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r02m.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02m_g1.json 2> gpurun_out/bench_r02m_g1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_r02m_g1.json") if l.startswith("{")][-1])
    print("g1", round(d["ms_per_step"],3), "host", d["host_enqueue_ms_per_step"], {k:round(v,3) for k,v in d["stages_ms"].items()})
    s=d["strong"]; print("strong", round(s["ms_per_step"],3), {k:round(v,3) for k,v in s["stages_ms"].items()})
    print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_all_copied"], "roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"])
except Exception as e: print("ERR", e)
PY
