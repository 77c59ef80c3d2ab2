#!/bin/bash
# final single-GPU validation of the round: build hook, GPU suite, smoke, both bench arms as the driver runs them
OUT=gpurun_out/r2final1; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest.log | cut -c1-250
cp gpurun_out/parity_report.json $OUT/ 2>/dev/null
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log | cut -c1-200
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "allocs", d["device_allocations_in_timed_window_rank0"])
print("roofline", d["roofline"]["frac"], d["roofline"]["second"]["issue_frac"], d["roofline"]["second"]["l1_data_pipe_frac"])
print("developed", d["developed"]["value"], "complete", d["complete_case"]["value"], "config4", d["config4"]["value_median_step"], d["config4"]["ms_per_step_rank0"], d["config4"]["device_allocations_per_step_rank0"])
print("config5", [(r["particles"], round(r["particles_per_s"]/1e9, 3)) for r in d["config5"]["rows"]])
PY
tail -3 $OUT/bench.err | cut -c1-200
