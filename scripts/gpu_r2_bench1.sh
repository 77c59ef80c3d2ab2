#!/bin/bash
OUT=gpurun_out/r2z; mkdir -p $OUT
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
for k in ("developed", "complete_case"):
    print(k, json.dumps(d.get(k))[:400])
PY
tail -3 $OUT/bench.err
