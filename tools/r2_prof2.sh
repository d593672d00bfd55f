#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in fft8 noexact noidx; do
OCL_SC_LIB=build_variants/lib$v.so ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_${v}_c2.csv -k regex:"k_(extent|momentum|deposit)" \
    python tools/prof_kick.py 1000000 63 4 >> gpurun_out/prof.log 2>&1
OCL_SC_LIB=build_variants/lib$v.so ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_${v}_c4.csv -k regex:"k_(extent|momentum|deposit)" \
    python tools/prof_kick.py 12500000 127 3 >> gpurun_out/prof.log 2>&1
done
for g in 0 1; do
  echo "== gk3 gather=$g" | tee -a gpurun_out/ab2.log
  OCL_SC_LIB=build_variants/libgk3.so OCL_SC_GATHER=$g timeout 300 python tools/time_kick_variants.py 1000000:63 12500000:127 2>&1 | tail -2 | tee -a gpurun_out/ab2.log
done
grep -h "k_extent" gpurun_out/r2b_*_c2.csv | tail -3
