#!/bin/bash
# development: build a complete variant of libcq_b200.so (every translation unit, -DCQ_PROFILING: the CQ_* tuning
# environment variables are read) without touching the shipped library
# usage: profiles/build_variant_full.sh NAME [-D...]   ->  build/variants/libcq_NAME.so  (run with CQ_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../customquad_b200"
mkdir -p ../build/variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -diag-suppress 177 -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DCQ_PROFILING"
for tu in cq_api cq_comm cq_launch_matrix cq_launch_vector cq_launch_scalar; do
  nvcc $F "$@" -c csrc/$tu.cu -o ../build/variants/${tu}_$name.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../build/variants/libcq_$name.so ../build/variants/cq_api_$name.o ../build/variants/cq_comm_$name.o ../build/variants/cq_launch_matrix_$name.o ../build/variants/cq_launch_vector_$name.o ../build/variants/cq_launch_scalar_$name.o -lcudart -ldl
echo build/variants/libcq_$name.so
