# round-2 third-session end measurement: GPU tests, smoke, bench (both arms), ncu launch list and full set of the same command
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2s3_pytest_gpu.txt; cat gpurun_out/r2s3_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s3_smoke.log 2>&1; tail -2 gpurun_out/r2s3_smoke.log
timeout 900 python bench.py > gpurun_out/r2s3_bench_n1.json 2> gpurun_out/r2s3_bench_n1.err; echo rc=$?; tail -2 gpurun_out/r2s3_bench_n1.err
timeout 400 python bench.py --impl reference > gpurun_out/r2s3_bench_reference.json 2> gpurun_out/r2s3_bench_reference.err; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s3_launches.csv python bench.py --no-e2e --no-cpu --no-extra --steps 2 --warmup 1 > gpurun_out/r2s3_ncu_launch.log 2>&1; echo rc=$?
timeout 400 ncu --set full --import-source on --clock-control none -k regex:assemble_moment_split -s 12 -c 4 -f -o gpurun_out/r2s3_prof_bench python bench.py --no-e2e --no-cpu --no-extra --steps 2 --warmup 1 > gpurun_out/r2s3_ncu_full.log 2>&1; echo rc=$?
ls -la gpurun_out/r2s3_prof_bench.ncu-rep
