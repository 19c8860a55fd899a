#!/bin/bash
# Builds A/B variants of the library into .variants/ (git-ignored; they travel to the GPU box with the snapshot):
#   scripts/build_variants.sh name "-DO2V_OCC_BATCH=32 -DO2V_OCC_THREADS=64" [name2 "flags2" ...]
# Run one with O2V_B200_LIB=.variants/libo2v_<name>.so python scripts/profile_run.py cfg4 4
set -e
cd "$(dirname "$0")/../obj2voxel_b200/csrc"
mkdir -p ../../.variants
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    dir=../../.variants/build_$name
    mkdir -p $dir
    for f in o2v_kernels o2v_sparse o2v_occupancy o2v_engine; do
        /usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --extended-lambda \
            -Xcompiler -fPIC,-ffp-contract=off -I../../include -I. $flags -c $f.cu -o $dir/$f.o &
    done
    wait
    for f in o2v_host_math o2v_capi o2v_io o2v_job o2v_bitscan; do
        g++ -std=c++17 -O2 -fPIC -ffp-contract=off -I../../include -I. -I/usr/local/cuda/include -c $f.cpp -o $dir/$f.o &
    done
    wait
    /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../.variants/libo2v_$name.so $dir/*.o -lcudart -lz -lpthread
    echo built .variants/libo2v_$name.so
done
