#!/bin/bash
# Build an A/B variant of the product library: scripts/build_variant.sh <name> "<extra nvcc flags>"
#   -> build/variants/lib_<name>.so  (selected at run time with RATILQR_B200_LIB=$PWD/build/variants/lib_<name>.so)
# Only the two translation units that hold the solve kernels are recompiled with the extra flags.
set -e
name=$1; shift
flags="$*"
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/ratilqr.jl_b200/csrc
out=$root/build/variants
mkdir -p $out
make -C $src -s -j8
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -I$src -diag-suppress 550,128"
$NV $flags -c $src/rl_kernels_solve.cu -o $out/solve_$name.o &
$NV $flags -c $src/rl_kernels_spec.cu -o $out/spec_$name.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/lib_$name.so $out/solve_$name.o $out/spec_$name.o $src/rl_kernels_coop.o $src/rl_kernels_comp.o $src/rl_capi.o $src/rl_user_host.o -ldl
echo built $out/lib_$name.so
