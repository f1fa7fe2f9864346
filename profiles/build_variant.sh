#!/bin/bash
# build_variant.sh <name> [extra nvcc flags]: an A/B build of the library -> build/ab/lib_<name>.so
name=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -split-compile=4 \
  -Xcompiler -fPIC -shared -diag-suppress 177 -I include "$@" -o build/ab/lib_$name.so sdaopt_b200/csrc/sda_kernels.cu
