set -x
# session 3, call 5: Q2 quadrilateral kernel with the zeroing warp and one expansion set (4 + 8 + 3 + 1 warps)
run() { CQ_LIB=$1 timeout 300 python bench.py --no-e2e --no-cpu --no-extra --steps 5 $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_form_ms'], [(k['kernel'], round(k['ms'],4), round(k['frac'],3)) for k in d['roofline']['per_kernel']])"; }
echo "== in-tree"; run customquad_b200/libcq_b200.so ""
echo "== in-tree fused"; run customquad_b200/libcq_b200.so --fused-zero
echo "== q83"; run build/variants/libcq_q83.so ""
echo "== fz83 fused"; run build/variants/libcq_fz83.so --fused-zero
echo "== fz73 fused"; run build/variants/libcq_fz73.so --fused-zero
for a in 4 10 16; do echo "== fz83 fused ahead $a"; CQ_ZERO_AHEAD=$a run build/variants/libcq_fz83.so --fused-zero; done
CQ_LIB=build/variants/libcq_fz83.so timeout 600 python -m pytest tests/test_gpu_fused_zero.py -q -m gpu -x 2>&1 | tail -2
