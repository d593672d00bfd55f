#!/bin/bash
# Run under gpurun: ncu evidence for the bench command (launch list) and the dominant kernel (full set).
# Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into the tracked files under profiles/.
set -u
TAG=${1:-r1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
for W in "1000000 63 c2" "12500000 127 c4"; do
  set -- $W
  ncu --set full --clock-control none --import-source on -k regex:"k_gather_kick" -s 2 -c 1 \
      -o gpurun_out/${TAG}_gather_kick_$3 -f python tools/prof_kick.py $1 $2 3 > gpurun_out/prof.log 2>&1
  ncu --set full --clock-control none -k regex:"k_(momentum|extent|deposit|cplx_outer|khat_z|rho_z|inv_z|field|green_table|real_even_outer)" -s 14 -c 14 \
      -o gpurun_out/${TAG}_all_kernels_$3 -f python tools/prof_kick.py $1 $2 3 >> gpurun_out/prof.log 2>&1
done
tail -2 gpurun_out/prof.log
