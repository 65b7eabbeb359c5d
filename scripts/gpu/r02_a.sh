#!/bin/bash
# round 2, GPU call A: parity tests + first measurements of the sync-free pipeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02a.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r02a.log
tail -15 gpurun_out/pytest_r02a.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/bench_r02a_l0.json 2> gpurun_out/bench_r02a_l0.err; echo "bench l0 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --no-e2e --pixel-layout 1 > gpurun_out/bench_r02a_l1.json 2> gpurun_out/bench_r02a_l1.err; echo "bench l1 rc=$?"
DDN_LIB_PATH=$PWD/depthdensifier_b200/libddn_b200_shfl.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong --no-e2e > gpurun_out/bench_r02a_shfl.json 2> gpurun_out/bench_r02a_shfl.err; echo "bench shfl rc=$?"
python - <<'PY'
import json
for f in ("l0","l1","shfl"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench_r02a_{f}.json") if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()}, d.get("e2e") and round(d["e2e"]["ms_per_step"],2))
    except Exception as e: print(f, "ERR", e)
PY
