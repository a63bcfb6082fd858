#!/bin/bash
# Builds build/libpib_exp.so: the library with the pair-kernel experiment variants (-DPIB_V6_VARIANTS plus any extra -D flags
# given as arguments).  Load it with PIB_LIB_PATH=build/libpib_exp.so (scripts/exp_pair_v6.py).
set -e
cd "$(dirname "$0")/../pyinterpolate_b200/csrc"
mkdir -p ../../build
OUT=${PIB_EXP_OUT:-libpib_exp.so}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v \
    -DPIB_V6_VARIANTS "$@" -c variogram.cu -o ../../build/variogram_exp.o 2> ../../build/variogram_exp.ptxas.log || { cat ../../build/variogram_exp.ptxas.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/$OUT ../../build/variogram_exp.o spatial.o kriging.o cabi.o -lcudart
