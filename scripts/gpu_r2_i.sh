#!/bin/bash
# round 2, call I (N GPUs): interior ranks (two neighbours) on the peer-mailbox rebuild + wall slabs: bit-identity check, then
# the bench at N (all sub-records when FULL=1).  usage: scripts/gpu_r2_i.sh <N> [FULL]
N=${1:-4}; FULL=${2:-0}; OUT=gpurun_out/r2i_n$N; mkdir -p $OUT
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
SPHB200_CHECK_EXCHANGE=1 run 150 29511 tests/multi_gpu_check.py --dp 0.025 --outer 30 --recut-interval 7 --cut-shift 2 --out $OUT/check.json > $OUT/check.log 2>&1
echo "check rc=$?"; grep MULTI_GPU_CHECK $OUT/check.log | head -1 | cut -c1-1000; tail -2 $OUT/check.log | cut -c1-300
if [ "$FULL" = "1" ]; then
  run 150 29513 tests/multi_gpu_check.py --dp 0.05 --outer 20 --recut-interval 7 --correction --surface-indicator --observers --out $OUT/complete_case.json > $OUT/complete_case.log 2>&1
  echo "complete case rc=$?"; grep MULTI_GPU_CHECK $OUT/complete_case.log | head -1 | cut -c1-700
fi
if [ "$FULL" = "1" ]; then ARGS="--steps 20 --warmup 5"; T=480; else ARGS="--steps 20 --warmup 5 --no-extras"; T=240; fi
run $T 29530 bench.py --gpus $N $ARGS > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "kernels", d["roofline"]["launch_ms"], d["roofline"]["other_kernels_ms"])
    for k in ("parity", "developed", "complete_case", "config3", "config4"):
        print(k, json.dumps(d.get(k))[:1100])
except Exception as e:
    print("no line:", e)
PY
tail -4 $OUT/bench.err | cut -c1-300
