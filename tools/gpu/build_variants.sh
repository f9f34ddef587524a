#!/bin/bash
# development aid: builds of the engine with experiment macros, for A/B timing on one box (EPH_B200_ENGINE_LIB)
cd "$(dirname "$0")/../../user-eph_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include --shared"
OUT=../../tools/gpu/variants
build() { name=$1; shift; nvcc $FLAGS "$@" -o $OUT/libeph_b200_$name.so eph_b200.cu eph_atomic.cu & }
"$@"
wait
