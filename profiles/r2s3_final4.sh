# final tree: the driver's sequence (GPU tests, smoke, bench N = 1)
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2s3_pytest_gpu_final.txt; cat gpurun_out/r2s3_pytest_gpu_final.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2s3_bench_n1d.json 2> gpurun_out/r2s3_bench_n1d.err; echo rc=$?; tail -1 gpurun_out/r2s3_bench_n1d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s3_bench_n1d.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step']['frac'], d['e2e']['ms_per_step'], d['gpu_launches'])
PY
