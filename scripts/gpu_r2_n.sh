#!/bin/bash
bash scripts/gpu_variants.sh r2n "-DSPH_REL_MINB=0" "-DSPH_REL_MINB=9" "-DSPH_REL_MINB=10" "-DSPH_REL_MINB=6"
