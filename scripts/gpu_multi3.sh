#!/bin/bash
# N-GPU validation: decomposition check (migration + re-cuts, strict bit-identity), then the bench at N. usage: scripts/gpu_multi3.sh <tag> <N> <bench-steps>
TAG=${1:-m}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29533 tests/multi_gpu_check.py --dp 0.025 --outer 30 --recut-interval 7 --cut-shift 2 --out $OUT/check.json > $OUT/check.log 2>&1; echo "check rc=$?"; grep MULTI_GPU_CHECK $OUT/check.log | head -1 | cut -c1-700; tail -2 $OUT/check.log | cut -c1-300
run 29535 bench.py --gpus $N --steps $3 --warmup 3 --no-cpu-baseline > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; tail -c 2600 $OUT/bench_n$N.json | cut -c1-300; tail -3 $OUT/bench_n$N.err | cut -c1-300
