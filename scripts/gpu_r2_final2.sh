#!/bin/bash
# final tree of round 2 (second session): GPU suite, smoke, the driver's bench command (both arms), launch list + full capture
OUT=gpurun_out/r2w; mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_ncu.log 2>&1; echo "launches rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k "regex:k_a1_interact|k_linear_correction|k_surface" -s 6 -c 6 -f -o $OUT/prof_cc \
    python scripts/complete_case_probe.py --outer 1 --warmup 1 > $OUT/prof_cc.log 2>&1; echo "full cc rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
for k in ("developed", "complete_case", "config4"):
    print(k, json.dumps(d.get(k))[:500])
print(open("$OUT/bench_reference.json").read()[:600])
PY
ls -la $OUT
