#!/bin/bash
# build a variant of libcnf_sm100.so with extra nvcc flags: scripts/build_variant.sh <name> [-DFLAG ...]
set -e
cd "$(dirname "$0")/../category_nerf_reconstruction_official_b200/csrc"
name=$1; shift
mkdir -p variants build/$name
for f in cnf_api cnf_sampler cnf_embed cnf_composite cnf_mlp_simt cnf_selftest; do cp build/$f.o build/$name/ 2>/dev/null || true; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c cnf_fused.cu -o build/$name/cnf_fused.o
nvcc -shared -o variants/libcnf_$name.so build/$name/*.o -cudart static
echo built variants/libcnf_$name.so
