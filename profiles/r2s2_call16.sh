set -x
( time timeout 2700 python -m pytest tests -q -m gpu 2>&1 | tail -8 ) 2>&1 | tail -12
