# development: Q2 forms with the in-kernel zeroing on, over variant libraries / environment settings
run() { # label, env...
  l=$1; shift
  env "$@" timeout 90 python bench.py --no-e2e --no-cpu --forms q2_stiffness,q2_mass 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$l'.ljust(28), round(d['ms_per_step'],3), [round(v,3) for v in d['per_form_ms'].values()])"
}
run "memset+kernel" CQ_ZERO_FUSED=0
run "fused a8" CQ_ZERO_FUSED=1 CQ_ZERO_AHEAD=8
run "fused a6" CQ_ZERO_FUSED=1 CQ_ZERO_AHEAD=6
run "fused a8 factor1.0" CQ_ZERO_FUSED=1 CQ_ZERO_AHEAD=8 CQ_STAGE_FACTOR=1.0
run "fused a8 sg8" CQ_ZERO_FUSED=1 CQ_ZERO_AHEAD=8 CQ_LIB=build/variants/libcq_sg8.so
run "fused a8 sg2" CQ_ZERO_FUSED=1 CQ_ZERO_AHEAD=8 CQ_LIB=build/variants/libcq_sg2.so
run "memset sg8" CQ_ZERO_FUSED=0 CQ_LIB=build/variants/libcq_sg8.so
run "fused a8 nmw7" CQ_ZERO_FUSED=1 CQ_ZERO_AHEAD=8 CQ_LIB=build/variants/libcq_nmw7.so
