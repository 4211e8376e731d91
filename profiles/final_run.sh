# round-end measurement: smoke, bench (both arms), ncu launch list and full set of the same command
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s3.log 2>&1; tail -2 gpurun_out/smoke_s3.log
timeout 400 python bench.py > gpurun_out/bench_s3_final.json 2> gpurun_out/bench_s3_final.err; tail -2 gpurun_out/bench_s3_final.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_s3_reference.json 2> gpurun_out/bench_s3_reference.err; tail -1 gpurun_out/bench_s3_reference.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s3.csv python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/ncu_launch_s3.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:assemble_moment_split -s 12 -c 4 -f -o gpurun_out/prof_s3 python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/ncu_full_s3.log 2>&1
ls -la gpurun_out/prof_s3.ncu-rep
