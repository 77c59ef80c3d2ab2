#!/bin/bash
# round 2, call F (1 GPU): GPU suite + config-4 leg at 256^3 with its kernel breakdown
OUT=gpurun_out/r2f; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
cp gpurun_out/parity_report.json $OUT/ 2>/dev/null
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --developed-steps 0 --config5-sizes 16 ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["roofline"]["other_kernels_ms"])
print("config4", json.dumps(d.get("config4")))
PY
