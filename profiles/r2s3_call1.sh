set -x
# session 3, call 1: long-rule shapes of the Q1 hexahedron moment kernel and SG = 8 lanes per entry on quadrilaterals
for lib in head base l3w10 l3w11; do echo "== $lib"
for r in "200 300" "100 120" "40 60"; do set -- $r; CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 96 1 stiffness $1 $2 5 | cut -c1-140; done
CQ_LIB=build/variants/libcq_$lib.so timeout 300 python bench.py --only-config 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_kernel_ms'], d['parity']['ok'])"
done
for lib in base sg8 base sg8; do echo "== $lib"
CQ_LIB=build/variants/libcq_$lib.so timeout 300 python bench.py --no-e2e --no-cpu --no-extra --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], [(k['kernel'], round(k['ms'],4), round(k['frac'],3)) for k in d['roofline']['per_kernel']])"
done
CQ_LIB=build/variants/libcq_sg8.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2
CQ_LIB=build/variants/libcq_l3w10.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "hexahedra" 2>&1 | tail -2
