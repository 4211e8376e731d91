# round-2 third-session end measurement, final tree: bench (both arms) + ncu launch list of the same command
set -x
timeout 900 python bench.py > gpurun_out/r2s3_bench_n1c.json 2> gpurun_out/r2s3_bench_n1c.err; echo rc=$?; tail -2 gpurun_out/r2s3_bench_n1c.err
timeout 400 python bench.py --impl reference > gpurun_out/r2s3_bench_referencec.json 2> gpurun_out/r2s3_bench_referencec.err; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s3_launches3.csv python bench.py --no-e2e --no-cpu --no-extra --steps 2 --warmup 1 > gpurun_out/r2s3_ncu_launch3.log 2>&1; echo rc=$?
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "device_resident or moment" 2>&1 | tail -2
