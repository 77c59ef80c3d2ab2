#!/bin/bash
# round 2, call B (1 GPU): full GPU suite, bench with the trimmed kernels, config-4 regression check, ncu launch list + full capture
OUT=gpurun_out/r2b; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
cp gpurun_out/parity_report.json $OUT/ 2>/dev/null
( time timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 1200 $OUT/bench.json
( time timeout 300 python scripts/config4_bench.py 256 4 ring ) > $OUT/config4.log 2>&1; echo "config4 rc=$?"; tail -3 $OUT/config4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/launches.log 2>&1; echo "launch list rc=$?"
timeout 900 bash scripts/gpu_prof.sh r2b "k_a2|k_a1_interact|k_compression_summation|k_relation_ordered" 20 8; echo "ncu full rc=$?"
