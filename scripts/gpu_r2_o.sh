#!/bin/bash
OUT=gpurun_out/r2o; mkdir -p $OUT
( time timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --developed-steps 0 --config5-sizes 16 ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "allocs in window", d["device_allocations_in_timed_window_rank0"])
c = d["config4"]; print("config4", c["ms_per_step"], c["ms_per_step_rank0"], c["device_allocations_per_step_rank0"], c["clocks"])
print("complete_case", json.dumps(d["complete_case"])[:600])
PY
tail -3 $OUT/bench.err | cut -c1-200
