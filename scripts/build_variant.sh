#!/bin/bash
# Tuning builds: scripts/build_variant.sh NAME FILE.cu "-DKNOB=value ..."  ->  scratch/libgrav_b200_NAME.so
# (FILE.cu recompiled with the extra defines, every other object taken from the regular build; select the result with
# GRAV_B200_LIB=scratch/libgrav_b200_NAME.so).  Not part of the product.
set -e
cd "$(dirname "$0")/../gravity-simulator_b200"
name=$1; file=$2; defs=$3
mkdir -p ../scratch build
base=$(basename "$file" .cu)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $defs -c csrc/$base.cu -o ../scratch/${base}_$name.o
objs=$(ls build/*.o | grep -v "build/$base.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../scratch/libgrav_b200_$name.so $objs ../scratch/${base}_$name.o -lnccl
echo "built scratch/libgrav_b200_$name.so"
