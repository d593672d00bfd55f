#!/bin/bash
# ncu atomic / reduction counters of the two deposit kernels
cd "$(dirname "$0")/.."
M="gpu__time_duration.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_requests_op_red.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed_op_global_red.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_op_red_hit_rate.pct"
for W in "1000000 63 c2" "12500000 127 c4" "50000000 255 c5"; do
  set -- $W
  ncu --metrics $M --clock-control none -k regex:"k_deposit" -s 2 -c 1 --csv --log-file gpurun_out/r2_atomics_deposit_$3.csv python tools/prof_kick.py $1 $2 3 > /dev/null 2>&1
done
ncu --metrics $M --clock-control none -k regex:"k_lsc_deposit" -s 2 -c 1 --csv --log-file gpurun_out/r2_atomics_lsc_deposit_c4.csv python tools/prof_lsc.py 12500000 > /dev/null 2>&1
grep -h "red\|duration" gpurun_out/r2_atomics_*.csv | awk -F'","' '{print $5" | "$(NF-2)" | "$NF}' | cut -c1-160
