#!/bin/bash
# Key figures of one ncu --set full capture: tools/ncu_summary.sh <file.ncu-rep>
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin))
h=r[0]; u=r[1]; v=r[2]
want=['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__waves_per_multiprocessor','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct','sm__inst_executed.sum','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__warps_eligible.avg.per_cycle_active','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__shared_mem_per_block_dynamic','launch__shared_mem_per_block_static','smsp__thread_inst_executed_per_inst_executed.ratio','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_inst0.avg.pct','sm__cycles_active.avg','smsp__warps_active.avg.per_cycle_active']
for i,n in enumerate(h):
    if n in want: print('%-62s %14s %s'%(n, v[i], u[i]))
out=[]
for i,n in enumerate(h):
    if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('per_issue_active.ratio') and 'not_issued' not in n:
        out.append((float(v[i].replace(',','')),n))
tot=sum(x for x,_ in out)
print('stall cycles per issued instruction (share of %.2f):'%tot)
for x,n in sorted(out,reverse=True)[:8]: print('  %6.2f %5.1f%%  %s'%(x,100*x/tot,n.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
"
