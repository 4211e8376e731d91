set -x
timeout 600 python -m pytest tests/test_gpu_owner.py -q -m gpu 2>&1 | grep -E "passed|failed|FAILED|Error" | cut -c1-200
for t in "False-3" "True-2" "True-3" "reproducible"; do timeout 300 python -m pytest tests/test_gpu_owner.py -q -m gpu -k "$t" 2>&1 | grep -E "passed|failed|FAILED|CqError" | cut -c1-200; done
