#!/bin/bash
# N GPUs: the complete case (correction record refreshed on the ghost planes) bit-identical to one GPU, then its timing
N=${1:-2}; OUT=gpurun_out/r2v_n$N; mkdir -p $OUT
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
SPHB200_CHECK_EXCHANGE=1 run 150 29513 tests/multi_gpu_check.py --dp 0.05 --outer 20 --recut-interval 7 --correction --surface-indicator --observers --out $OUT/complete_case.json > $OUT/complete_case.log 2>&1
echo "complete case rc=$?"; grep MULTI_GPU_CHECK $OUT/complete_case.log | head -1 | cut -c1-900
SPHB200_CHECK_EXCHANGE=1 run 150 29515 tests/multi_gpu_check.py --dp 0.025 --outer 30 --recut-interval 7 --cut-shift 2 --correction --out $OUT/check_corr.json > $OUT/check_corr.log 2>&1
echo "correction dp=0.025 rc=$?"; grep MULTI_GPU_CHECK $OUT/check_corr.log | head -1 | cut -c1-900
