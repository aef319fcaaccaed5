#!/bin/bash
# Build a variant of libcnf_sm100.so whose cnf_fused.cu comes from another source directory (development aid for same-box A/B timing):
#   scripts/build_variant.sh <name> <dir with cnf_fused.cu cnf_fused_a.inc cnf_fused_b.inc cnf_umma.cuh> [extra nvcc flags]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CS=$ROOT/category_nerf_reconstruction_official_b200/csrc
name=$1; src=$2; shift; shift
work=/tmp/variant_$name/pkg/csrc
rm -rf /tmp/variant_$name; mkdir -p $work /tmp/variant_$name/include $CS/variants $CS/build/$name
cp $ROOT/include/cnf.h /tmp/variant_$name/include/
cp $CS/cnf_common.cuh $work/
for f in cnf_fused.cu cnf_fused_a.inc cnf_fused_b.inc cnf_fused_c.inc cnf_umma.cuh cnf_sample_math.cuh; do cp $src/$f $work/; done
(cd $work && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c cnf_fused.cu -o cnf_fused.o 2>&1 | grep -A2 "fused_step_kernelILb0" | grep -E "spill" | sed "s/^/[$name] /")
for f in cnf_api cnf_sampler cnf_embed cnf_composite cnf_mlp_simt cnf_occmap_tc; do cp $CS/build/$f.o $CS/build/$name/; done
cp $work/cnf_fused.o $CS/build/$name/
nvcc -shared -o $CS/variants/libcnf_$name.so $CS/build/$name/*.o -cudart static 2>/dev/null
echo "built variants/libcnf_$name.so"
