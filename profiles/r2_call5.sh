# round 2, call 5 (2 GPUs): multi-GPU parity worker, bench at N=2 with the config-5 leg
set -x
nvidia-smi topo -m 2>/dev/null | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r2_pytest5.log 2>&1; tail -15 gpurun_out/r2_pytest5.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench5_n2.json 2> gpurun_out/r2_bench5_n2.err; echo rc=$?; tail -25 gpurun_out/r2_bench5_n2.err
