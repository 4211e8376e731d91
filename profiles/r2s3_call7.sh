set -x
# session 3, call 7: hexahedron vector / scalar kernels with 8 instead of 12 consumer warps (168 registers: no spills); coefficient loads issued early
for lib in customquad_b200/libcq_b200.so build/variants/libcq_vs38.so; do echo "== $lib"
for c in "hex 96 1 load 8 8" "hex 96 1 load 8 27" "hex 96 1 load 100 120" "hex 96 1 load_coef 8 27" "hex 96 1 nitsche_rhs 4 16" "hex 96 1 nitsche_rhs 8 27" "hex 96 1 scalar 8 27" "quad 1024 1 load_coef 16 64" "quad 1024 2 load_coef 16 64"; do
CQ_LIB=$lib python profiles/one_form.py $c 5 | cut -c1-140
done
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"assemble_staged" -s 2 -c 1 -o gpurun_out/r2s3_hex_nitsche_rhs -f python profiles/one_form.py hex 96 1 nitsche_rhs 4 16 2 > gpurun_out/r2s3_ncu_nrhs.log 2>&1; echo rc=$?
CQ_LIB=build/variants/libcq_vs38.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2
