set -x
timeout 600 python -m pytest tests/test_gpu_owner.py tests/test_gpu_parity.py -x -q -m gpu -k "owner or overwrite or reproducible or 3-2" 2>&1 | tail -3
CQ_BS=3 CQ_ZERO=1 python profiles/one_form.py hex 48 2 stiffness 27 27 5
CQ_BS=1 CQ_ZERO=0 python profiles/one_form.py hex 48 2 stiffness 27 27 5
timeout 600 python examples/vector_poisson.py -config4 128 > gpurun_out/r2s2_config4_owner4.json 2> gpurun_out/r2s2_config4_owner4.err; echo rc=$?; tail -3 gpurun_out/r2s2_config4_owner4.err; cat gpurun_out/r2s2_config4_owner4.json
bash profiles/hexsweep.sh customquad_b200/libcq_b200.so build/variants/libcq_npp2_6.so build/variants/libcq_npp2_4.so build/variants/libcq_npp3_4.so
