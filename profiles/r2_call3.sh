# round 2, call 3b: entry-major stage header (rowstart / mu) + lanes-per-entry variants; parity first
set -x
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest3.log 2>&1; tail -5 gpurun_out/r2_pytest3.log
for l in customquad_b200/libcq_b200.so; do
 for fz in 0 1; do
  CQ_ZERO_FUSED=$fz CQ_LIB=$l timeout 120 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$l fz=$fz', round(d['ms_per_step'],3), [round(v,3) for v in d['per_form_ms'].values()], [(round(k['ms'],3),round(k['frac'],3)) for k in d['roofline']['per_kernel']])"
 done
done
for l in customquad_b200/libcq_b200.so build/variants/libcq_g3_4.so build/variants/libcq_g3_8.so; do
 echo "== $l"
 CQ_LIB=$l timeout 200 python profiles/run_configs.py hexq1 2>&1 | python -c "
import sys,json
for x in sys.stdin:
    try: d=json.loads(x)
    except:
        if 'rror' in x: print(x.strip()[:200])
        continue
    print('  ', d['case'][:34].ljust(34), d['form'].ljust(10), '%.3f ms'%d['ms'], 'frac %.3f'%d['frac_hbm'])"
done
