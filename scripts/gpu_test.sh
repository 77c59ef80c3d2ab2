#!/bin/bash
# quick GPU round: parity tests (+ optional short bench). usage: scripts/gpu_test.sh <tag> [bench-steps]
TAG=${1:-t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > $OUT/pytest_gpu.log; cat $OUT/pytest_gpu.log | tail -25
if [ -n "$2" ]; then python bench.py --steps $2 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 2500 $OUT/bench.json; tail -5 $OUT/bench.err; fi
