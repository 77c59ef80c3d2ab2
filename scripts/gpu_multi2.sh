#!/bin/bash
# N-GPU round with and without exchange overlap. usage: scripts/gpu_multi2.sh <tag> <N> <bench-steps>
TAG=${1:-m}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29533 tests/multi_gpu_check.py --dp 0.05 --outer 12 --out $OUT/check_dp0.05.json > $OUT/check.log 2>&1; echo "check rc=$?"; grep MULTI_GPU_CHECK $OUT/check.log | cut -c1-600; tail -3 $OUT/check.log
run 29534 tests/multi_gpu_check.py --dp 0.025 --outer 20 --out $OUT/check_dp0.025.json > $OUT/check2.log 2>&1; echo "check2 rc=$?"; grep MULTI_GPU_CHECK $OUT/check2.log | cut -c1-600; tail -3 $OUT/check2.log
run 29535 bench.py --gpus $N --steps $3 --warmup 3 --no-cpu-baseline > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; tail -c 1500 $OUT/bench_n$N.json | cut -c1-400; tail -3 $OUT/bench_n$N.err
run 29536 bench.py --gpus $N --steps $3 --warmup 3 --no-cpu-baseline --serial-exchange > $OUT/bench_serial_n$N.json 2> $OUT/bench_serial_n$N.err; echo "bench(serial) rc=$?"; tail -c 1500 $OUT/bench_serial_n$N.json | cut -c1-400
