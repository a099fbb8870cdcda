#!/bin/bash
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/e2e_breakdown.py 2>&1 | grep -E "ms \(|Error|error" | head -20
