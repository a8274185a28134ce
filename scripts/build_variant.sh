#!/bin/bash
# dev tool: build libmntr_gpu_<name>.so with extra -D flags on linear_single.cu (the other objects are reused)
# usage: scripts/build_variant.sh name -DMNTR_K1_GROUP=6 ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC -Xptxas -v "$@" \
  -c minotaur_b200/csrc/linear_single.cu -o minotaur_b200/build/linear_single_$name.o 2> minotaur_b200/build/linear_single_$name.log
grep -A2 "DirectedELb1" minotaur_b200/build/linear_single_$name.log | grep -E "spill|Used" | sed "s/^/[$name] /"
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o minotaur_b200/libmntr_gpu_$name.so minotaur_b200/build/linear_single_$name.o \
  minotaur_b200/build/linear_batch.o minotaur_b200/build/linear_rounds.o minotaur_b200/build/mntr_gpu.o \
  -Xlinker --no-as-needed -lstdc++ -lm -ldl -lpthread -lrt
