#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it (all legs)
N=${1:-2}; OUT=gpurun_out/r2y_n$N; mkdir -p $OUT
timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
    for k in ("parity", "complete_case", "config4"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print("no line:", e)
PY
tail -3 $OUT/bench.err | cut -c1-300
