set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hexahedra or 3-1" 2>&1 | tail -3
bash profiles/hexsweep.sh customquad_b200/libcq_b200.so
python profiles/run_configs.py hexq1 2>/dev/null | cut -c1-160
CQ_BS=3 CQ_ZERO=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"assemble_staged" -s 2 -c 1 -o gpurun_out/r2s2_q2hex_phase1b -f python profiles/one_form.py hex 48 2 stiffness 27 27 2 > gpurun_out/r2s2_ncu_p1b.log 2>&1; echo rc=$?
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"assemble_moment_split" -s 2 -c 1 -o gpurun_out/r2s2_hexq1_stiff8_rot -f python profiles/one_form.py hex 128 1 stiffness 8 8 2 > gpurun_out/r2s2_ncu_hex.log 2>&1; echo rc=$?
