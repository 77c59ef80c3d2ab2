#!/bin/bash
# round 2, call C (1 GPU): A/B of the pair-arithmetic trims and occupancy settings of the interaction kernels
bash scripts/gpu_variants.sh r2c "-DSPH_TRIM=0" "-DSPH_TRIM=7" "-DSPH_TRIM=1" "-DSPH_TRIM=2" "-DSPH_TRIM=4" "-DSPH_TRIM=6" \
   "-DSPH_TRIM=7 -DSPH_FL_MIN_BLOCKS=7" "-DSPH_TRIM=2 -DSPH_FL_MIN_BLOCKS=7" "-DSPH_TRIM=0 -DSPH_FL_MIN_BLOCKS=6"
