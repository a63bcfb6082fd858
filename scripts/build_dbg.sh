#!/bin/bash
# Builds build/libpib_dbg.so: the whole library with -DPIB_BOUNDS_CHECK (index / shared-memory offset checks in the hot kernels,
# common.cuh).  Run the GPU parity tests against it:  PIB_LIB_PATH=build/libpib_dbg.so python -m pytest tests -m gpu -q
set -e
cd "$(dirname "$0")/../pyinterpolate_b200/csrc"
mkdir -p ../../build/dbg
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -DPIB_BOUNDS_CHECK"
for f in spatial variogram kriging cabi; do
    nvcc $FLAGS -c $f.cu -o ../../build/dbg/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/libpib_dbg.so ../../build/dbg/spatial.o ../../build/dbg/variogram.o \
    ../../build/dbg/kriging.o ../../build/dbg/cabi.o -lcudart
