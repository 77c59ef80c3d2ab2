#!/bin/bash
# round 2, call A (1 GPU): full GPU test suite (incl. the new full-size / variant parity tests), bench both arms
OUT=gpurun_out/r2a; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
cp gpurun_out/parity_report.json $OUT/ 2>/dev/null
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
( time timeout 400 python bench.py --impl reference --steps 20 --warmup 5 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -c 1500 $OUT/bench_ref.json
