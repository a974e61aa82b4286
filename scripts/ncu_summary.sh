#!/bin/bash
# usage: scripts/ncu_summary.sh file.ncu-rep  -> key counters per captured launch
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rd=list(csv.reader(sys.stdin))
hdr=rd[0]; units=rd[1]; rows=rd[2:]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','lts__t_bytes.sum','lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__cycles_active.avg','sm__cycles_elapsed.avg','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum']
idx={h:i for i,h in enumerate(hdr)}
for r in rows:
  print('----')
  for w in want:
    if w in idx: print(f'{w:75s} {r[idx[w]][:90]:>30s} {units[idx[w]]}')
"
