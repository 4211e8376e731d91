set -x
timeout 600 python -m pytest tests/test_gpu_levelset_device.py tests/test_gpu_cutfem_poisson.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python profiles/config3_real_rules.py 2>&1 | tail -8
timeout 120 python bench.py --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', round(d['ms_per_step'],3), [(round(k['ms'],3),round(k['frac'],3)) for k in d['roofline']['per_kernel']])"
