#!/bin/bash
# dev: leaf insert variants on the 16K^2 terrain (CPVS_LEAF_BATCH = leaves per thread of the staged insert; 0 = one leaf per thread, no staging)
for b in 0 1 2 4; do echo "== CPVS_LEAF_BATCH=$b"; CPVS_LEAF_BATCH=$b python scripts/one_build.py 16384 terrain 4 2>&1 | tail -1; done
