set -x
# session 3, N = 2: multi-GPU worker under pytest, the new device-coefficient test, bench --gpus 2 (weak-scaled headline + config 5 leg + interface parity)
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "device_resident" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r2s3_pytest_multi_gpu.txt; cat gpurun_out/r2s3_pytest_multi_gpu.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2s3_bench_n2.json 2> gpurun_out/r2s3_bench_n2.err; echo rc=$?
tail -4 gpurun_out/r2s3_bench_n2.err
