#!/bin/bash
OUT=gpurun_out/r2m; mkdir -p $OUT
timeout 300 python scripts/e2e_probe.py > $OUT/e2e_probe.log 2>&1; echo "rc=$?"; grep E2E_PROBE $OUT/e2e_probe.log; tail -3 $OUT/e2e_probe.log | cut -c1-300
