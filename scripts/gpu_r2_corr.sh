#!/bin/bash
# correction path with the 32-byte B record: full GPU suite, complete-case timing + launch list, base kernel times
OUT=gpurun_out/r2u; mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
python scripts/complete_case_probe.py 2>>$OUT/cc.err | grep COMPLETE_CASE | tee -a $OUT/complete_case.txt
python scripts/kbench.py --tag="corr_record_tree" 2>>$OUT/kbench.err | tee -a $OUT/variants.jsonl
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/cc_launches.csv \
    python scripts/complete_case_probe.py --outer 2 --warmup 1 > $OUT/cc_ncu.log 2>&1; echo "cc launches rc=$?"
cp gpurun_out/parity_report.json $OUT/ 2>/dev/null
ls $OUT
