# full GPU test suite (timed), smoke, ncu launch list + full set of the bench command
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 ) 2>&1 | tail -9
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s2_launches.csv python bench.py --no-e2e --no-cpu --no-extra --steps 2 --warmup 1 > gpurun_out/r2s2_ncu_launch.log 2>&1; echo rc=$?
timeout 400 ncu --set full --import-source on --clock-control none -k regex:assemble_moment_split -s 12 -c 4 -f -o gpurun_out/r2s2_prof_bench python bench.py --no-e2e --no-cpu --no-extra --steps 2 --warmup 1 > gpurun_out/r2s2_ncu_full.log 2>&1; echo rc=$?
ls -la gpurun_out/r2s2_prof_bench.ncu-rep
