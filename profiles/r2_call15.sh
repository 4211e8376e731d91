set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_exact_element_tensors.py tests/test_gpu_levelset_device.py -m gpu -q -x 2>&1 | tail -3
timeout 120 python bench.py --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', round(d['ms_per_step'],3), [(round(k['ms'],3),round(k['frac'],3)) for k in d['roofline']['per_kernel']])"
timeout 300 python profiles/run_configs.py 2>/dev/null | python -c "
import sys,json
for x in sys.stdin:
    try: d=json.loads(x)
    except: continue
    print('  ', d['case'][:34].ljust(34), d['form'].ljust(10), '%.3f ms'%d['ms'], 'frac %.3f'%d['frac_hbm'])"
timeout 300 python profiles/config5_single.py 256 2>&1 | grep -v Warning | head -3
timeout 600 python profiles/config3_real_rules.py 2>&1 | tail -8
