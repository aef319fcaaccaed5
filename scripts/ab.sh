#!/bin/bash
# same-box A/B timing of variant libraries: scripts/ab.sh A B C ...   (two interleaved rounds)
V=category_nerf_reconstruction_official_b200/csrc/variants
for round in 1 2; do for v in "$@"; do CNF_LIB_PATH=$V/libcnf_$v.so python scripts/time_fused.py $v; done; done
