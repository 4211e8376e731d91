set -x
timeout 600 python -m pytest tests/test_gpu_owner.py -x -q -m gpu 2>&1 | tail -60 | cut -c1-300
