#!/bin/bash
# build libsphb200 with several tuning-macro sets ON THE GPU BOX is not possible (no need: nvcc is there too) -- so: for each
# variant rebuild fluid.o with the macros, relink, run scripts/kbench.py. usage: scripts/gpu_variants.sh <tag> "<macros 1>" "<macros 2>" ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
cd sphinxsys_b200/csrc
for V in "$@"; do
  rm -f build/fluid.o
  case "$V" in *SPH_REL*) rm -f build/neighbor.o;; esac
  make NVCCFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-extended-lambda -ccbin /usr/bin/g++ $V" > /dev/null 2>&1 || echo "build failed: $V"
  (cd ../..; python scripts/kbench.py --tag="$V" | tee -a $OUT/variants.jsonl)
done
rm -f build/fluid.o build/neighbor.o; make > /dev/null 2>&1
