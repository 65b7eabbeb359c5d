#!/bin/bash
mkdir -p gpurun_out
KREGEX='regex:align_stats|remap_median|build_pair|bbox_init|backproject_filter|grid_from|grid_store|clear_units|mark_|tile_count|tile_scan|zero_accum|unit_prefix|accumulate_|finalize_kernel|merge_|unmark'
timeout 900 python -m pytest tests/test_gpu_session.py -m gpu -q > gpurun_out/pytest_r02e.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r02e.log
timeout 600 ncu -k "$KREGEX" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/merge_launches_r02e.csv python scripts/gpu/merge_bench.py cfg3 8 > gpurun_out/ncu_merge_r02e.log 2>&1; echo "ncu merge rc=$?"; tail -2 gpurun_out/ncu_merge_r02e.log | cut -c1-900
timeout 900 ncu -k "$KREGEX" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r02e.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-e2e > gpurun_out/ncu_r02e.log 2>&1; echo "ncu bench rc=$?"
python scripts/summarise_launches.py gpurun_out/launches_r02e.csv | tail -16
timeout 900 ncu -k "$KREGEX" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3_25v_r02e.csv python scripts/gpu/kbench.py cfg3_25 3 > gpurun_out/ncu_cfg3_25_r02e.log 2>&1; echo "ncu cfg3_25 rc=$?"; tail -1 gpurun_out/ncu_cfg3_25_r02e.log | cut -c1-900
