#!/bin/bash
# development: vector / scalar form timings (profiles/run_configs.py hexq1 quadnitsche) for a list of variant libraries
for v in "$@"; do
  echo "== $v"
  CQ_LIB=$v timeout 150 python profiles/run_configs.py hexq1 quadnitsche 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    print('  ', d['case'][:34].ljust(34), d['form'].ljust(10), '%.3f ms'%d['ms'], 'frac %.3f'%d['frac_hbm'])"
done
