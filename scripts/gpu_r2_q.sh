#!/bin/bash
# stage trace of a decomposed run on an interior rank (two neighbours)
N=${1:-4}; OUT=gpurun_out/r2q_n$N; mkdir -p $OUT
SPHB200_STEP_TRACE=1 TRACE_RANK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 scripts/long_run_probe.py 0.003937 30 30 > $OUT/trace.log 2>&1
echo "rc=$?"; grep -A30 STEP_TRACE $OUT/trace.log | cut -c1-180
