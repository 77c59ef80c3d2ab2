#!/bin/bash
# round 2, call H (1 GPU): config-4 step at 256^3 — plain, without the ring position refresh, with the stage trace
OUT=gpurun_out/r2h; mkdir -p $OUT
timeout 300 python scripts/config4_bench.py 256 4 ring > $OUT/plain.log 2>&1; echo "plain rc=$?"; grep "^ring" $OUT/plain.log | cut -c1-200
SPHB200_NO_RING_POSITION_REFRESH=1 timeout 300 python scripts/config4_bench.py 256 4 ring > $OUT/norefresh.log 2>&1; echo "no refresh rc=$?"; grep "^ring" $OUT/norefresh.log | cut -c1-200
timeout 300 python scripts/config4_bench.py 256 4 ring > $OUT/plain2.log 2>&1; echo "plain again rc=$?"; grep "^ring" $OUT/plain2.log | cut -c1-200
timeout 300 python scripts/config4_bench.py 256 12 ring > $OUT/plain12.log 2>&1; echo "plain 12 steps rc=$?"; grep "^ring" $OUT/plain12.log | cut -c1-200
