set -x
# session 3, call 3: points in flight of the quadrilateral moment loop, long-rule shape on quadrilaterals, deeper unroll of the long hexahedron shape, three expansion sets in the hexahedron Nitsche kernel
for lib in base q2u3 q2u4; do echo "== $lib"
CQ_LIB=build/variants/libcq_$lib.so timeout 300 python bench.py --no-e2e --no-cpu --no-extra --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], [(k['kernel'], round(k['ms'],4), round(k['frac'],3)) for k in d['roofline']['per_kernel']])"
done
for lib in base q2long q2longu4; do echo "== $lib"
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py quad 1024 2 stiffness 200 300 5 | cut -c1-140
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py quad 1024 2 mass 80 120 5 | cut -c1-140
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py quad 1024 1 stiffness 200 300 5 | cut -c1-140
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py quad 1024 1 stiffness 80 120 5 | cut -c1-140
done
for lib in base u6 u8; do echo "== $lib"
for r in "200 300" "100 120" "40 60"; do set -- $r; CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 96 1 stiffness $1 $2 5 | cut -c1-140; done
done
for lib in nitx3; do echo "== $lib"
for r in "4 16" "8 27" "20 30"; do set -- $r; CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 96 1 nitsche $1 $2 5 | cut -c1-140; done
done
timeout 300 python bench.py --only-config 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_kernel_ms'], d['roofline']['frac'], d['parity']['ok'])"
CQ_LIB=build/variants/libcq_nitx3.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "hexahedra" 2>&1 | tail -2
CQ_LIB=build/variants/libcq_q2longu4.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "moment or overflow" 2>&1 | tail -2
