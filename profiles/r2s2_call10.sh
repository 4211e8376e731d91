set -x
timeout 600 python -m pytest tests/test_gpu_owner.py -q -m gpu 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-200
