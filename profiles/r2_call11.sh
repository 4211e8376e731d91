set -x
timeout 900 python examples/vector_poisson.py -config4 128 > gpurun_out/r2_config4_128.json 2> gpurun_out/r2_config4_128.err; tail -3 gpurun_out/r2_config4_128.err; cat gpurun_out/r2_config4_128.json
CQ_L2_TILE_BYTES=1000000000000 timeout 900 python examples/vector_poisson.py -config4 128 2>/dev/null | tail -1
timeout 300 python profiles/run_configs.py hexq2 2>&1 | grep -v Warn
