set -x
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2s2_bench_n8.json 2> gpurun_out/r2s2_bench_n8.err; echo rc=$?
tail -6 gpurun_out/r2s2_bench_n8.err; cut -c1-600 gpurun_out/r2s2_bench_n8.json
