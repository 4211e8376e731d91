set -x
python profiles/one_form.py hex 96 1 stiffness 100 120 5
python profiles/one_form.py hex 96 1 stiffness 40 60 5
python profiles/one_form.py hex 96 1 stiffness 20 30 5
python profiles/one_form.py hex 96 1 nitsche 20 30 5
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"assemble_moment_split" -s 2 -c 1 -o gpurun_out/r2s2_hexq1_long -f python profiles/one_form.py hex 96 1 stiffness 100 120 2 > gpurun_out/r2s2_ncu_hexlong.log 2>&1; echo rc=$?
