set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest10.log 2>&1; tail -6 gpurun_out/r2_pytest10.log
timeout 900 python examples/vector_poisson.py -config4 128 > gpurun_out/r2_config4_128.json 2> gpurun_out/r2_config4_128.err; tail -3 gpurun_out/r2_config4_128.err; cat gpurun_out/r2_config4_128.json
timeout 300 python examples/vector_poisson.py -N 32 2>&1 | tail -1
timeout 300 python examples/vector_poisson.py -N 64 2>&1 | tail -1
