set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_exact_element_tensors.py tests/test_gpu_full_size.py -m gpu -q -x -k "not config2 and not long_list" 2>&1 | tail -4
for l in customquad_b200/libcq_b200.so build/variants/libcq_dmb8.so build/variants/libcq_dmb10.so build/variants/libcq_dmb13.so; do
 echo "== $l"
 CQ_LIB=$l timeout 200 python profiles/run_configs.py hexq2 2>/dev/null | python -c "
import sys,json
for x in sys.stdin:
    try: d=json.loads(x)
    except: continue
    print('  ', d['case'][:34].ljust(34), d['form'].ljust(10), '%.3f ms'%d['ms'], 'frac %.3f'%d['frac_hbm'])"
done
timeout 900 python examples/vector_poisson.py -config4 128 2>/dev/null | tail -1
