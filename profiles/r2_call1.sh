# round 2, call 1: sanity (tests), bench baseline, L2-prefetch A/B on bench + other config shapes, fused zero A/B
set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; tail -3 gpurun_out/r2_pytest1.log
timeout 400 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -2 gpurun_out/r2_bench1.err
for l in customquad_b200/libcq_b200.so build/variants/libcq_nopf.so; do
 for fz in 0 1; do
  CQ_ZERO_FUSED=$fz CQ_LIB=$l timeout 120 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$l fz=$fz', round(d['ms_per_step'],3), [round(v,3) for v in d['per_form_ms'].values()], [(round(k['ms'],3),round(k['frac'],3)) for k in d['roofline']['per_kernel']])"
 done
 echo "== $l"
 CQ_LIB=$l timeout 200 python profiles/run_configs.py 2>/dev/null | python -c "
import sys,json
for x in sys.stdin:
    try: d=json.loads(x)
    except: continue
    print('  ', d['case'][:34].ljust(34), d['form'].ljust(10), '%.3f ms'%d['ms'], 'frac %.3f'%d['frac_hbm'])"
done
