#!/bin/bash
# development: hex Q1 timings (profiles/run_configs.py hexq1) for a list of variant libraries
for v in "$@"; do
  echo "== $v"
  CQ_LIB=$v timeout 120 python profiles/run_configs.py hexq1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    if d['form'] in ('stiffness','mass'): print('  ', d['case'][:30].ljust(30), d['form'].ljust(10), '%.3f ms'%d['ms'], 'frac %.3f'%d['frac_hbm'])"
done
