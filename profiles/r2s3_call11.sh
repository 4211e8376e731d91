set -x
# session 3, N = 2: zero_keep protocol (interface rows zeroed before the fence, the rest inside the call) -- multi-GPU worker and the weak-scaled headline with interface parity
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r2s3_pytest_multi_gpu2.txt; cat gpurun_out/r2s3_pytest_multi_gpu2.txt
CQ_BENCH_C5=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2s3_bench_n2b.json 2> gpurun_out/r2s3_bench_n2b.err; echo rc=$?
tail -3 gpurun_out/r2s3_bench_n2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s3_bench_n2b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['per_form_ms'], {k:v for k,v in d['parity'].items() if k!='what'})
PY
