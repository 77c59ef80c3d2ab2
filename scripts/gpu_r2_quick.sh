#!/bin/bash
# quick validation of the tree: GPU suite, smoke, complete-case timing, base kernel times
OUT=gpurun_out/${1:-r2x}; mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
python scripts/complete_case_probe.py 2>>$OUT/cc.err | grep COMPLETE_CASE | tee -a $OUT/complete_case.txt
python scripts/kbench.py --tag="${2:-tree}" 2>>$OUT/kbench.err | tee -a $OUT/variants.jsonl
cp gpurun_out/parity_report.json $OUT/ 2>/dev/null
