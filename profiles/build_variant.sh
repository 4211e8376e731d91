#!/bin/bash
# development: build a variant of libcq_b200.so with extra -D flags for cq_launch_matrix.cu
# usage: profiles/build_variant.sh NAME -DCQ_MOMENT_NMW=12 ...   ->  build/variants/libcq_NAME.so  (run with CQ_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../customquad_b200"
mkdir -p ../build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -diag-suppress 177 -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DCQ_PROFILING "$@" -c csrc/cq_launch_matrix.cu -o ../build/variants/lm_$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../build/variants/libcq_$name.so csrc/cq_api.o csrc/cq_comm.o ../build/variants/lm_$name.o csrc/cq_launch_vector.o csrc/cq_launch_scalar.o -lcudart -ldl
echo build/variants/libcq_$name.so
