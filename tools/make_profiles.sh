#!/bin/bash
# Run under gpurun: ncu evidence for the bench command (launch list) and the dominant kernel (full set).
# Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into the tracked files under profiles/.
# The .ncu-rep files of the all-kernel captures are condensed on the box (gpurun copies back at most
# 64 MiB): PARTS selects what to capture, e.g. PARTS="bench gather all lsc c5".
set -u
TAG=${1:-r1}
PARTS=${PARTS:-"bench gather all lsc c5"}
has() { [[ " $PARTS " == *" $1 "* ]]; }
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
has bench && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
for W in "1000000 63 c2" "12500000 127 c4"; do
  set -- $W
  has gather && ncu --set full --clock-control none --import-source on -k regex:"k_gather_kick" -s 2 -c 1 \
      -o gpurun_out/${TAG}_gather_kick_$3 -f python tools/prof_kick.py $1 $2 3 > gpurun_out/prof.log 2>&1
  has all && ncu --set full --clock-control none -k regex:"k_(momentum|extent|deposit|cplx_outer|khat_z|rho_z|inv_z|field|green_table|real_even_outer)" -s 14 -c 14 \
      -o gpurun_out/${TAG}_all_kernels_$3 -f python tools/prof_kick.py $1 $2 3 >> gpurun_out/prof.log 2>&1
done
# longitudinal space charge: every kernel of one kick (12.5 M particles, the third kick)
has lsc && ncu --set full --clock-control none -k regex:"k_lsc_" -s 17 -c 8 \
    -o gpurun_out/${TAG}_lsc_kernels_c4 -f python tools/prof_lsc.py 12500000 >> gpurun_out/prof.log 2>&1
# 255^3 mesh (512^3 box): launch list of one kick
has c5 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_c5.csv \
    python tools/prof_kick.py 50000000 255 3 >> gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
# condense on the box, keep only the dominant kernel's full reports
PROFILES_OUT=gpurun_out/profiles_${TAG} python tools/summarise_profiles.py ${TAG}
rm -f gpurun_out/${TAG}_all_kernels_*.ncu-rep gpurun_out/${TAG}_lsc_kernels_*.ncu-rep
