#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_session.py -m gpu -q > gpurun_out/pytest_r02p.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r02p.log
KREGEX='regex:merge_|tile_scan|zero_accum|finalize_kernel|unmark'
timeout 600 ncu -k "$KREGEX" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/merge_launches_r02p.csv python scripts/gpu/merge_bench.py cfg3 8 > gpurun_out/ncu_merge_r02p.log 2>&1; echo "ncu merge rc=$?"; tail -1 gpurun_out/ncu_merge_r02p.log | cut -c1-600
python scripts/summarise_launches.py gpurun_out/merge_launches_r02p.csv merge_gather_prefix | tail -11
