#!/bin/bash
# usage: scripts/ncu_launches.sh <tag> [bench args...]   (run on the GPU box through gpurun)
# Launch list of this library's kernels for one bench step: duration + DRAM bytes per launch.
tag=$1; shift
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"mark_|tile_|unit_prefix|zero_accum|accumulate_|finalize_|align_stats|remap_median|backproject|build_pair|voxel_key|segment_mean|merge_segments|DeviceRadixSort|DeviceScan|DeviceRunLength" \
  -c 80 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_${tag}.log 2>&1
tail -c 200 gpurun_out/ncu_${tag}.log
