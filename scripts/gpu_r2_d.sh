#!/bin/bash
# round 2, call D (N GPUs): peer-mailbox rebuild against the NCCL rebuild — bit-identity with one GPU (migration, re-cuts,
# complete case), then the bench at N with the parity digest and the pipelined e2e, both spellings.  usage: scripts/gpu_r2_d.sh <N>
N=${1:-2}; OUT=gpurun_out/r2d_n$N; mkdir -p $OUT
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for MODE in 1 0; do
  export SPHB200_PEER_REBUILD=$MODE
  SPHB200_CHECK_EXCHANGE=1 run 300 2951$MODE tests/multi_gpu_check.py --dp 0.025 --outer 30 --recut-interval 7 --cut-shift 2 --out $OUT/check_peer$MODE.json > $OUT/check_peer$MODE.log 2>&1
  echo "check peer=$MODE rc=$?"; grep MULTI_GPU_CHECK $OUT/check_peer$MODE.log | head -1 | cut -c1-600; tail -2 $OUT/check_peer$MODE.log | cut -c1-300
done
export SPHB200_PEER_REBUILD=1
run 300 29520 tests/multi_gpu_check.py --dp 0.05 --outer 20 --recut-interval 7 --correction --surface-indicator --observers --out $OUT/complete_case.json > $OUT/complete_case.log 2>&1
echo "complete case rc=$?"; grep MULTI_GPU_CHECK $OUT/complete_case.log | head -1 | cut -c1-600
for MODE in 1 0; do
  export SPHB200_PEER_REBUILD=$MODE
  run 600 2953$MODE bench.py --gpus $N --steps 20 --warmup 5 --no-extras > $OUT/bench_peer$MODE.json 2> $OUT/bench_peer$MODE.err
  echo "bench peer=$MODE rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_peer$MODE.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "rebuild ms", d["roofline"]["other_kernels_ms"]["cell_list_build+reorder"])
    print("parity", json.dumps(d["parity"])[:900])
except Exception as e:
    print("no line:", e)
PY
  tail -3 $OUT/bench_peer$MODE.err | cut -c1-300
done
