#!/bin/bash
# round 2, call J (2 GPUs): long horizon — bit-identity over 600 advection steps with re-cuts at the case cadence, and the
# course of the time-step ratio / max velocity at the resolution of the 8-GPU bench on one and on two GPUs
OUT=gpurun_out/r2j; mkdir -p $OUT
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 300 29511 tests/multi_gpu_check.py --dp 0.0125 --outer 600 --recut-interval 100 --out $OUT/check_long.json > $OUT/check_long.log 2>&1
echo "long check rc=$?"; grep MULTI_GPU_CHECK $OUT/check_long.log | head -1 | cut -c1-900; tail -2 $OUT/check_long.log | cut -c1-300
CUDA_VISIBLE_DEVICES=0 timeout 400 python scripts/long_run_probe.py 0.003125 1300 100 > $OUT/probe_1gpu.log 2>&1 &
P1=$!
sleep 1
wait $P1; echo "probe 1gpu rc=$?"; grep PROBE $OUT/probe_1gpu.log | cut -c1-260
run 300 29520 scripts/long_run_probe.py 0.00496 1300 100 > $OUT/probe_2gpu.log 2>&1; echo "probe 2gpu rc=$?"; grep PROBE $OUT/probe_2gpu.log | cut -c1-300
