#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in 0 1; do
OCL_SC_GATHER=$g ncu --set full --clock-control none --import-source on -k regex:"k_gather_kick" -s 2 -c 1 \
   -o gpurun_out/r2a_gather_g${g}_c4 -f python tools/prof_kick.py 12500000 127 3 > gpurun_out/prof.log 2>&1
done
OCL_SC_LIB=build_variants/libfft8.so ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_c2.csv \
    python tools/prof_kick.py 1000000 63 4 >> gpurun_out/prof.log 2>&1
OCL_SC_LIB=build_variants/libfft8.so ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_c4.csv \
    python tools/prof_kick.py 12500000 127 3 >> gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
ls -la gpurun_out/*.ncu-rep
