set -x
./profiles/microbench/row_store_bench
timeout 600 python -m pytest tests/test_gpu_owner.py -q -m gpu 2>&1 | tail -2
timeout 600 python examples/vector_poisson.py -config4 128 2>/dev/null | cut -c1-700
