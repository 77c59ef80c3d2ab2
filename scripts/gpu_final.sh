#!/bin/bash
# Final-tree evidence in one short call: ncu launch list of the bench command, then a full capture of the hot kernels.
TAG=${1:-v8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_ncu.log 2>&1; echo "launches rc=$?"
timeout 110 ncu --set full --clock-control none --import-source on -k "regex:k_a2|k_a1_interact|k_relation_ordered|k_compression" -s 40 -c 6 -f -o $OUT/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/prof.log 2>&1; echo "full rc=$?"
ls -la $OUT
