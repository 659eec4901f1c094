#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]   ->  build/variants/<name>.so
# (a copy of the product library compiled with other -D switches, for A/B timing:
#  PINTS_B200_LIB=build/variants/<name>.so python tools/br_time.py)
NAME=$1; shift
mkdir -p build/variants
nvcc -O3 -std=c++17 -lineinfo -shared --threads 4 -Xcompiler -fPIC -Xcompiler -O3 -cudart static \
    -gencode arch=compute_100a,code=sm_100a "$@" \
    pints_b200/csrc/pb200_api.cu pints_b200/csrc/pb200_ode.cu pints_b200/csrc/pb200_analytic.cu \
    pints_b200/csrc/pb200_samplers.cu pints_b200/csrc/pb200_rank.cu pints_b200/csrc/pb200_exchange.cu \
    -o build/variants/$NAME.so
