set -x
export CQ_LIB=build/variants/libcq_pipe.so
timeout 300 python -m pytest tests/test_gpu_owner.py -q -m gpu 2>&1 | tail -2
for cfg in "1 80" "16 80" "16 64" "16 96" "8 80" "32 80"; do set -- $cfg; echo "== chunks $1 sms $2"; CQ_OWNER_CHUNKS=$1 CQ_OWNER_SMS=$2 timeout 300 python examples/vector_poisson.py -config4 128 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('stiffness_zero_fused_ms','stiffness_ms','stiffness_total_sum_rel','mass_total_rel_err')})"; done
CQ_OWNER_CHUNKS=16 CQ_OWNER_SMS=80 timeout 600 python bench.py --only-config 4 2>/dev/null | cut -c1-900
