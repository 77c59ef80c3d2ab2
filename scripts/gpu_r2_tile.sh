#!/bin/bash
# (needs profiles/r02b_tile_order_experiment.patch applied: SPHB200_TILE_ORDER is not in the tree)
# SM-local tile order A/B: bit-identity of the state, kernel times, full test suite with the order forced on, launch list of the complete case
OUT=gpurun_out/r2t; mkdir -p $OUT
for T in 0 1; do SPHB200_TILE_ORDER=$T python scripts/state_hash.py 0.0125 2>&1 | grep STATE_HASH | sed "s/^/tile=$T /" | tee -a $OUT/hash.txt; done
for T in 0 1; do SPHB200_TILE_ORDER=$T python scripts/kbench.py --tag="tile_order=$T" 2>>$OUT/kbench.err | tee -a $OUT/variants.jsonl; done
for T in 0 1; do SPHB200_TILE_ORDER=$T python scripts/complete_case_probe.py 2>>$OUT/cc.err | grep COMPLETE_CASE | tee -a $OUT/complete_case.txt; done
SPHB200_TILE_ORDER=1 SPHB200_TILE_MIN=0 timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_tile1.log 2>&1; echo "pytest(tile=1,min=0) rc=$?"; tail -3 $OUT/pytest_tile1.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/cc_launches.csv \
    python scripts/complete_case_probe.py --outer 2 --warmup 1 > $OUT/cc_ncu.log 2>&1; echo "cc launches rc=$?"
SPHB200_TILE_ORDER=1 timeout 150 ncu --set full --clock-control none -k "regex:k_a2|k_a1_interact|k_compression" -s 30 -c 12 -f -o $OUT/prof_tile1 \
    python scripts/kbench.py --outer 1 > $OUT/prof.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
