#!/bin/bash
# usage: scripts/ncu_keymetrics.sh report.ncu-rep   -> the metrics the design notes quote
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin))
hdr,units,vals=r[0],r[1],r[2]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_atom.sum','lts__t_sectors_op_red.sum','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_xu.sum','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.sum','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed']
for i,h in enumerate(hdr):
    if h in want or h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_warp_active.pct')==False and 'ratio' in h:
        try:
            v=float(vals[i].replace(',',''))
            if h.startswith('smsp__average_warp') and v<0.15: continue
        except: pass
        print(f'{h:90s} {units[i]:12s} {vals[i]}')
"
