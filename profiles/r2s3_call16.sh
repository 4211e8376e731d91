set -x
# session 3, call 16: final tree (6 + 6 + 4 short-rule hexahedron shape, Nitsche loop one point in flight): parity, fused-zero tests, config 3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_zero.py tests/test_gpu_cutfem_poisson.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python bench.py --only-config 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_kernel_ms'], d['roofline']['frac'], d['parity']['ok'])"
