#!/bin/bash
# overlap of stage 1 with the previous step: N GPUs (N from $1), with and without
mkdir -p gpurun_out
N=${1:-1}
run() { if [ "$N" = "1" ]; then timeout 900 python bench.py "$@"; else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; fi; }
if [ "$N" = "1" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_session.py -m gpu -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02s_pytest.log; fi
run --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02s_overlap_g$N.json 2> gpurun_out/r02s_overlap_g$N.err; echo "overlap rc=$?"
run --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-overlap-align --no-check > gpurun_out/r02s_nooverlap_g$N.json 2> gpurun_out/r02s_nooverlap_g$N.err; echo "no overlap rc=$?"
python - <<PY
import json
for f in ("overlap", "nooverlap"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02s_{f}_g$N.json") if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), "strong", round(d["strong"]["ms_per_step"],3), d.get("multi_gpu_check") and d["multi_gpu_check"]["passed"], {k:round(v,3) for k,v in d["strong"]["stages_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
