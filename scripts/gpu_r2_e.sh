#!/bin/bash
# round 2, call E (N GPUs): slabs of the wall — bit-identity with one GPU incl. re-cuts that force the wall subset to be
# reloaded, then the complete case.  usage: scripts/gpu_r2_e.sh <N>
N=${1:-2}; OUT=gpurun_out/r2e_n$N; mkdir -p $OUT
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
SPHB200_CHECK_EXCHANGE=1 run 200 29511 tests/multi_gpu_check.py --dp 0.025 --outer 40 --recut-interval 5 --cut-shift 6 --out $OUT/check_wall_slab.json > $OUT/check_wall_slab.log 2>&1
echo "wall slab check rc=$?"; grep MULTI_GPU_CHECK $OUT/check_wall_slab.log | head -1 | cut -c1-900; tail -3 $OUT/check_wall_slab.log | cut -c1-300
run 300 29520 tests/multi_gpu_check.py --dp 0.05 --outer 20 --recut-interval 7 --correction --surface-indicator --observers --out $OUT/complete_case.json > $OUT/complete_case.log 2>&1
echo "complete case rc=$?"; grep MULTI_GPU_CHECK $OUT/complete_case.log | head -1 | cut -c1-600; tail -3 $OUT/complete_case.log | cut -c1-300
# the bench at N with every sub-record, config 3 forced at a reduced size (leg debugging at small cost)
run 900 29530 bench.py --gpus $N --steps 10 --warmup 3 --developed-steps 100 --force-config3 --config3-dp ${CONFIG3_DP:-0.004} --config4-side ${CONFIG4_SIDE:-128} > $OUT/bench_full.json 2> $OUT/bench_full.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_full.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "parity", "developed", "config3", "config4", "e2e"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print("no line:", e)
PY
tail -5 $OUT/bench_full.err | cut -c1-300
