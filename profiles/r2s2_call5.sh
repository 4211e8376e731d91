set -x
timeout 600 python -m pytest tests/test_gpu_owner.py -x -q -m gpu 2>&1 | tail -3
CQ_BS=3 CQ_ZERO=1 python profiles/one_form.py hex 48 2 stiffness 27 27 5
timeout 600 python examples/vector_poisson.py -config4 128 > gpurun_out/r2s2_config4_owner3.json 2> gpurun_out/r2s2_config4_owner3.err; echo rc=$?; tail -3 gpurun_out/r2s2_config4_owner3.err; cat gpurun_out/r2s2_config4_owner3.json
CQ_BS=3 CQ_ZERO=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"assemble_staged" -s 2 -c 1 -o gpurun_out/r2s2_q2hex_phase1 -f python profiles/one_form.py hex 48 2 stiffness 27 27 2 > gpurun_out/r2s2_ncu_p1.log 2>&1; echo rc=$?
CQ_BS=3 CQ_ZERO=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"owner_rows" -s 2 -c 1 -o gpurun_out/r2s2_q2hex_phase2 -f python profiles/one_form.py hex 48 2 stiffness 27 27 2 > gpurun_out/r2s2_ncu_p2.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
