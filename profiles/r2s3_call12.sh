set -x
# session 3, call 12: look-ahead of the zeroing warp with the wait in the moment warp; hexahedron kernels with the zeroing warp again (wait in the moment warp); parity on the final tree
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_zero.py -q -m gpu -x 2>&1 | tail -2
run() { CQ_LIB=build/variants/libcq_cur.so timeout 300 python bench.py --no-e2e --no-cpu --no-extra --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_form_ms'])"; }
for a in 6 9 12; do echo "== ahead $a"; CQ_ZERO_AHEAD=$a run; done
for f in 0 1; do echo "== hex fused $f"
CQ_ZERO=1 CQ_FUSED=$f python profiles/one_form.py hex 128 1 stiffness 8 8 7 | cut -c1-140
CQ_ZERO=1 CQ_FUSED=$f python profiles/one_form.py hex 128 1 stiffness 8 27 7 | cut -c1-140
done
