#!/bin/bash
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29600
for m in 0 8; do for an in 0 1; do
  port=$((port+1)); echo -n "NVLS_MODE=$m announce=$an "; PROBE_MODE=nvls OCL_SC_NVLS_ANNOUNCE=$an OCL_SC_NVLS_MODE=$m $TR --master-port $port tools/r2_comm_probe.py 2>&1 | grep -E "graph|Error|error" | tail -2
done; done
port=$((port+1)); echo -n "skip=4  "; PROBE_MODE=nvls OCL_SC_DEBUG_SKIP=4 $TR --master-port $port tools/r2_comm_probe.py 2>&1 | grep -E "graph|Error|error" | tail -2
