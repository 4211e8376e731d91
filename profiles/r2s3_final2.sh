# round-2 third-session end measurement (after the zeroing policy change): fused-zero + parity tests, ncu full set of the step's kernels -> traffic.json, bench (both arms), ncu launch list
set -x
timeout 900 python -m pytest tests/test_gpu_fused_zero.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 400 ncu --set full --import-source on --clock-control none -k regex:assemble_moment_split -s 12 -c 4 -f -o gpurun_out/r2s3_prof_bench2 python bench.py --no-e2e --no-cpu --no-extra --steps 2 --warmup 1 > gpurun_out/r2s3_ncu_full2.log 2>&1; echo rc=$?
ncu -i gpurun_out/r2s3_prof_bench2.ncu-rep --page raw --csv > gpurun_out/r2s3_bench_raw2.csv 2>/dev/null
python profiles/make_traffic.py gpurun_out/r2s3_bench_raw2.csv profiles/traffic.json; cp profiles/traffic.json gpurun_out/r2s3_traffic.json
timeout 900 python bench.py > gpurun_out/r2s3_bench_n1b.json 2> gpurun_out/r2s3_bench_n1b.err; echo rc=$?; tail -2 gpurun_out/r2s3_bench_n1b.err
timeout 400 python bench.py --impl reference > gpurun_out/r2s3_bench_referenceb.json 2> gpurun_out/r2s3_bench_referenceb.err; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s3_launches2.csv python bench.py --no-e2e --no-cpu --no-extra --steps 2 --warmup 1 > gpurun_out/r2s3_ncu_launch2.log 2>&1; echo rc=$?
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
