#!/bin/bash
# round 2, GPU call B: all parity tests (no -x), stage micro-benchmark, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02b.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r02b.log
tail -25 gpurun_out/pytest_r02b.log
timeout 600 python scripts/gpu/kbench.py cfg2 7 > gpurun_out/kbench_r02b.json 2> gpurun_out/kbench_r02b.err; echo "kbench rc=$?"; cat gpurun_out/kbench_r02b.json; tail -3 gpurun_out/kbench_r02b.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-e2e > gpurun_out/ncu_r02b.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r02b.json") if l.startswith("{")][-1])
print(round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()}, d.get("e2e") and (round(d["e2e"]["ms_per_step"],2), round(d["e2e"]["ms_per_step_all_copied"],2)))
PY
python scripts/summarise_launches.py gpurun_out/launches_r02b.csv 2>/dev/null | tail -30
