set -x
# session 3, call 2: in-kernel zeroEntries on the Q1 hexahedron moment kernels; unroll 4 of the hexahedron point loop; moment warps of the Nitsche kernel
timeout 900 python -m pytest tests/test_gpu_fused_zero.py -q -m gpu -x 2>&1 | tail -3
for f in 0 1; do echo "== fused $f"
for r in "8 8" "8 27"; do set -- $r
CQ_ZERO=1 CQ_FUSED=$f python profiles/one_form.py hex 128 1 stiffness $1 $2 7 | cut -c1-140
done
CQ_ZERO=1 CQ_FUSED=$f python profiles/one_form.py hex 128 1 mass 8 8 7 | cut -c1-140
done
for fl in "--no-fused-zero" ""; do echo "== config 3 $fl"
timeout 300 python bench.py --only-config 3 $fl 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_kernel_ms'], d['roofline']['frac'], d['parity']['ok'])"
done
for lib in u4; do echo "== $lib"
for r in "200 300" "100 120" "40 60" "8 27"; do set -- $r; CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 96 1 stiffness $1 $2 5 | cut -c1-140; done
done
for lib in base nit8 nit10 nit11; do echo "== $lib"
for r in "4 16" "8 27" "20 30"; do set -- $r; CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 96 1 nitsche $1 $2 5 | cut -c1-140; done
done
CQ_LIB=build/variants/libcq_nit10.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "hexahedra" 2>&1 | tail -2
