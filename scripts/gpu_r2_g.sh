#!/bin/bash
# round 2, call G (1 GPU): where the config-4 step goes (stage trace), GPU suite, prefetch variants
OUT=gpurun_out/r2g; mkdir -p $OUT
SPHB200_STEP_TRACE=1 timeout 300 python scripts/config4_bench.py 256 3 ring > $OUT/config4_trace.log 2>&1; echo "trace rc=$?"; grep -A25 STEP_TRACE $OUT/config4_trace.log | cut -c1-200
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest.log | cut -c1-200
cp gpurun_out/parity_report.json $OUT/ 2>/dev/null
bash scripts/gpu_variants.sh r2g "-DSPH_PREFETCH=0" "-DSPH_PREFETCH=1" "-DSPH_PREFETCH=2"
