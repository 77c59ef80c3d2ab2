#!/bin/bash
# ncu --set full of selected kernels only. usage: scripts/gpu_prof.sh <tag> <kernel-regex> [skip] [count]
TAG=${1:-p}; KRE=${2:-k_a2}; SKIP=${3:-20}; CNT=${4:-4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $CNT -f -o $OUT/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/prof.log 2>&1
ls -la $OUT
