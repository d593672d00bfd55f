#!/bin/bash
# 2-GPU: cost of each cross-rank synchronisation of the sharded kick (OCL_SC_DEBUG_SKIP leaves parts out; timing only)
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29600
for skip in 0 1 2 4 7; do
  port=$((port+1)); echo -n "skip=$skip  "; PROBE_MODE=nvls OCL_SC_DEBUG_SKIP=$skip $TR --master-port $port tools/r2_comm_probe.py 2>&1 | grep -E "graph|Error|error" | tail -2
done
