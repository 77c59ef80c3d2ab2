#!/bin/bash
# round 2, call L (1 GPU): ncu launch list + full capture of the hot kernels on the current tree, then the bench line
OUT=gpurun_out/r2l; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/launches.log 2>&1; echo "launch list rc=$?"
timeout 900 bash scripts/gpu_prof.sh r2l "k_a2|k_a1_interact|k_compression_summation|k_relation_ordered|k_cell_count|k_gather_multi" 20 12; echo "ncu full rc=$?"
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 600 $OUT/bench.json
