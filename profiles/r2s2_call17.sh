set -x
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2s2_bench_n2.json 2> gpurun_out/r2s2_bench_n2.err; echo rc=$?
tail -8 gpurun_out/r2s2_bench_n2.err; cut -c1-3000 gpurun_out/r2s2_bench_n2.json
