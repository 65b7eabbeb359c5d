#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu/kbench.py cfg2 5 > gpurun_out/kbench_r02d.json 2> gpurun_out/kbench_r02d.err; cat gpurun_out/kbench_r02d.json; tail -2 gpurun_out/kbench_r02d.err
timeout 600 python scripts/gpu/merge_bench.py cfg3 2 8 > gpurun_out/merge_bench_r02d.json 2> gpurun_out/merge_bench_r02d.err; cat gpurun_out/merge_bench_r02d.json; tail -3 gpurun_out/merge_bench_r02d.err
timeout 900 ncu -k regex:'merge_|tile_scan|zero_accum|finalize' --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/merge_launches_r02d.csv python scripts/gpu/merge_bench.py cfg3 8 > gpurun_out/ncu_merge_r02d.log 2>&1; echo "ncu merge rc=$?"
timeout 900 ncu -k regex:'ddn' --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r02d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-e2e > gpurun_out/ncu_r02d.log 2>&1; echo "ncu bench rc=$?"
python scripts/summarise_launches.py gpurun_out/launches_r02d.csv | tail -16
