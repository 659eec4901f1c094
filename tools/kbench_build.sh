#!/bin/bash
# usage: tools/kbench_build.sh <out> [extra nvcc flags...]
OUT=$1; shift
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --threads 4 "$@" \
    tools/kbench.cu pints_b200/csrc/pb200_api.cu pints_b200/csrc/pb200_ode.cu \
    pints_b200/csrc/pb200_analytic.cu pints_b200/csrc/pb200_samplers.cu \
    pints_b200/csrc/pb200_rank.cu pints_b200/csrc/pb200_exchange.cu -o $OUT
