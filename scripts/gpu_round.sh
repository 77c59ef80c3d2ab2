#!/bin/bash
# One gpurun call: GPU parity tests, a short bench, an ncu launch list and full captures of the hot kernels.
# usage: scripts/gpu_round.sh <tag> [kernel-regex]
TAG=${1:-run}
KRE=${2:-'k_a2|k_a1_interact|k_relation|k_compression'}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s 30 -c 12 -f -o $OUT/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/prof.log 2>&1
ls -la $OUT
