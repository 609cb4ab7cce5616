#!/bin/bash
# Builds an experimental variant of the engine next to the product library (never loaded unless CKKS_B200_LIB points at it):
#   profiles/build_variant.sh NAME -DV_FLAG [-DV_OTHER=2 ...]   ->  profiles/variants/libckks_NAME.so
set -e
name="$1"; shift
cd "$(dirname "$0")/../seal-fyp-logistic-regression_b200/csrc"
mkdir -p ../../profiles/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -o ../../profiles/variants/libckks_$name.so engine.cu tables.cpp
echo built profiles/variants/libckks_$name.so
