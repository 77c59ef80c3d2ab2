#!/bin/bash
# ring-of-one GPU test first (new code), then the rest of the GPU suite
OUT=gpurun_out/ring; mkdir -p $OUT
timeout 70 python -m pytest tests/test_gpu_zz_periodic_ring.py -q -m gpu 2>&1 | tail -60 > $OUT/pytest_ring.log; echo "ring rc=${PIPESTATUS[0]}" >> $OUT/pytest_ring.log
timeout 60 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_zz_periodic_ring.py 2>&1 | tail -15 > $OUT/pytest_all.log
tail -5 $OUT/pytest_ring.log; tail -3 $OUT/pytest_all.log
