set -x
timeout 1200 python bench.py > gpurun_out/r2s2_bench_n1.json 2> gpurun_out/r2s2_bench_n1.err; echo rc=$?
tail -25 gpurun_out/r2s2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s2_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','per_form_ms')})
print('e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'])
for k in ('config3','config4'):
    print(k, json.dumps(d.get('extra',{}).get(k))[:1800])
print(d.get('extra',{}).get('error'))
PY
