#!/bin/bash
# Development helper: build a variant of the library with extra -D flags for rollout_tc.cu into build/variants/<name>.so
#   scripts/build_variant.sh <name> -DDP_ISSUERS=0,1,2,3,4
# and time it with   DP_LIB_VARIANT=<name> python scripts/time_rollout.py tc
set -e
cd "$(dirname "$0")/../density_planner_b200/csrc"
name=$1; shift
mkdir -p ../../build/variants
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v -ccbin /usr/bin/g++ "$@" -c rollout_tc.cu -o ../../build/variants/$name.o 2> ../../build/variants/$name.ptxas.log
objs="api.o step.o cost.o bin.o planner.o pipeline.o dnn.o"   # the product objects (make them first)
$NVCC $ARCH -shared -o ../../build/variants/$name.so $objs ../../build/variants/$name.o -ccbin /usr/bin/g++ 2>&1 | tail -3
grep -A1 "ILi1ELi1ELi0ELi0ELi1" ../../build/variants/$name.ptxas.log | grep spill
