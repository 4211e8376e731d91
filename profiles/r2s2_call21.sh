set -x
# the coordinate magnitudes of ranks 6 / 7 of the 8-GPU weak-scaling mesh on two GPUs: x in [14.85, 20.17]
CQ_BENCH_XSHIFT=15.96 CQ_BENCH_C5=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2s2_bench_n2_shift.json 2> gpurun_out/r2s2_bench_n2_shift.err; echo rc=$?
tail -3 gpurun_out/r2s2_bench_n2_shift.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s2_bench_n2_shift.json').read().strip().splitlines()[-1])
print({k:v for k,v in d['parity'].items() if k!='what'})
PY
