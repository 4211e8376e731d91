#!/bin/bash
# development: build a variant of libcq_b200.so with extra -D flags for ALL launch translation units
# usage: profiles/build_variant_all.sh NAME -DCQ_STAGED_NPW_VS=8 ...   ->  build/variants/libcq_NAME.so  (run with CQ_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../customquad_b200"
mkdir -p ../build/variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -diag-suppress 177 -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DCQ_PROFILING"
for tu in vector scalar; do
  nvcc $F "$@" -c csrc/cq_launch_$tu.cu -o ../build/variants/l${tu}_$name.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../build/variants/libcq_$name.so csrc/cq_api.o csrc/cq_comm.o csrc/cq_launch_matrix.o ../build/variants/lvector_$name.o ../build/variants/lscalar_$name.o -lcudart -ldl
echo build/variants/libcq_$name.so
