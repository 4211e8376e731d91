set -x
for dyn in 0 1; do for tile in 25165824 1000000000000; do
echo "== config4 128^3 dyn=$dyn tile=$tile"
CQ_DYNAMIC=$dyn CQ_L2_TILE_BYTES=$tile CQ_LIB=build/variants/libcq_prof.so timeout 900 python examples/vector_poisson.py -config4 128 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['stiffness_ms'], d['mass_ms'], d['stiffness_frac_hbm_6450'], d['stiffness_total_sum_rel'], d['mass_total_rel_err'])"
done; done
