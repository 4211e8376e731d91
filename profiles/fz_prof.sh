# development: ncu full set + source counters of the fused-zeroing Q2 mass kernel
CQ_ZERO_AHEAD=${AHEAD:-8} timeout 300 ncu --set full --import-source on --clock-control none -k regex:assemble_moment_split -s 3 -c 1 -f -o gpurun_out/prof_fz python bench.py --no-e2e --no-cpu --steps 1 --forms q2_mass > gpurun_out/ncu_prof_fz.log 2>&1
ls -la gpurun_out/prof_fz.ncu-rep
