#!/bin/bash
# round-2 A/B run: GPU tests, then stage timings for library variants x gather layouts
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
for lib in "" build_variants/libfft8.so build_variants/libfft16.so; do
  for g in 1 0; do
    echo "== lib=${lib:-default} gather=$g" | tee -a gpurun_out/ab.log
    OCL_SC_LIB=$lib OCL_SC_GATHER=$g timeout 300 python tools/time_kick_variants.py 200000:31 1000000:63 12500000:127 2>&1 | tail -4 | tee -a gpurun_out/ab.log
  done
done
