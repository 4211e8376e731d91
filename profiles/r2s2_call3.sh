set -x
timeout 900 python -m pytest tests/test_gpu_owner.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu -k "owner or overwrite or reproducible or 3-2 or config4 or element_tensors" 2>&1 | tail -15
python profiles/run_configs.py hexq2 2>&1 | cut -c1-220
timeout 600 python examples/vector_poisson.py -config4 128 > gpurun_out/r2s2_config4_owner2.json 2> gpurun_out/r2s2_config4_owner2.err; echo rc=$?; tail -3 gpurun_out/r2s2_config4_owner2.err; cat gpurun_out/r2s2_config4_owner2.json
