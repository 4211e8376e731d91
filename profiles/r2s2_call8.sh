set -x
timeout 800 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_owner.py -x -q -m gpu -k "False-2" 2>&1 | grep -v "^$" | head -60 | cut -c1-250
