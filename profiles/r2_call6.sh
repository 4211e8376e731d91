# round 2, call 6 (2 GPUs): traversal-tile hint: parity test, then config 5 at N=2 with and without it
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "plane_tile or hexahedra" > gpurun_out/r2_pytest6.log 2>&1; tail -5 gpurun_out/r2_pytest6.log
for tile in 25165824 1000000000000; do
CQ_L2_TILE_BYTES=$tile timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2_bench6_n2_$tile.json 2> gpurun_out/r2_bench6_n2_$tile.err; echo rc=$?; tail -3 gpurun_out/r2_bench6_n2_$tile.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench6_n2_$tile.json')); c=d['extra']['config5']; print('tile=$tile', c.get('ms_per_step'), c.get('per_gpu_roofline',{}).get('frac'), c.get('parity'), c.get('nccl_exchange_path'), c.get('error'))"
done
