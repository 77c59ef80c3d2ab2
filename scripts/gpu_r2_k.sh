#!/bin/bash
# round 2, call K (N GPUs): e2e path with packed host vectors and NUMA-local pinned buffers
N=${1:-2}; OUT=gpurun_out/r2k_n$N; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-parity > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"]); print("e2e", json.dumps(d["e2e"])[:1200])
except Exception as e:
    print("no line:", e)
PY
tail -4 $OUT/bench.err | cut -c1-300
numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -14
