set -x
timeout 900 python -m pytest tests/test_gpu_owner.py tests/test_gpu_parity.py -x -q -m gpu -k "owner or overwrite or reproducible or 3-2-3 or far_from" 2>&1 | tail -15
python profiles/run_configs.py hexq2 2>&1 | cut -c1-220
timeout 600 python examples/vector_poisson.py -config4 128 > gpurun_out/r2s2_config4_owner.json 2> gpurun_out/r2s2_config4_owner.err; echo rc=$?; tail -3 gpurun_out/r2s2_config4_owner.err; cat gpurun_out/r2s2_config4_owner.json
