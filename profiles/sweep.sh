#!/bin/bash
# development sweep: per-kernel ms of the bench workload under environment switches
# usage: profiles/sweep.sh "VAR=a VAR2=b" ...   (one bench run per argument)
for cfg in "$@"; do
  env $cfg python bench.py --no-e2e --no-cpu --steps 3 ${SWEEP_FORMS:+--forms $SWEEP_FORMS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg', ' '.join('%s=%.3f' % (k['kernel'], k['ms']) for k in d['roofline']['per_kernel']), 'step=%.3f' % d['ms_per_step'])"
done
