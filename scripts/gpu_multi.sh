#!/bin/bash
# N-GPU round: decomposition parity check, then bench at N. usage: scripts/gpu_multi.sh <tag> <N> [bench-steps]
TAG=${1:-m}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tests/multi_gpu_check.py --dp 0.05 --outer 12 --out $OUT/check_dp0.05.json > $OUT/check.log 2>&1; echo "check rc=$?"; grep MULTI_GPU_CHECK $OUT/check.log; tail -5 $OUT/check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    tests/multi_gpu_check.py --dp 0.025 --outer 20 --out $OUT/check_dp0.025.json > $OUT/check2.log 2>&1; echo "check2 rc=$?"; grep MULTI_GPU_CHECK $OUT/check2.log; tail -3 $OUT/check2.log
if [ -n "$3" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 \
      bench.py --gpus $N --steps $3 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; tail -c 2500 $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
fi
