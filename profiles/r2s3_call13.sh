set -x
# session 3, call 13: zeroing by the producer warps (ZW = 2: no zeroing warp, the kernels keep their shapes)
export CQ_LIB=build/variants/libcq_zp.so
timeout 400 python -m pytest tests/test_gpu_fused_zero.py -q -m gpu -x 2>&1 | tail -3
run() { timeout 200 python bench.py --no-e2e --no-cpu --no-extra --steps 5 $1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_form_ms'])"; }
echo "== zp default"; run ""
echo "== zp fused all"; run --fused-zero
echo "== zp memset"; run --no-fused-zero
for f in 0 1; do echo "== hex fused $f"
CQ_ZERO=1 CQ_FUSED=$f timeout 120 python profiles/one_form.py hex 128 1 stiffness 8 8 7 | cut -c1-140
CQ_ZERO=1 CQ_FUSED=$f timeout 120 python profiles/one_form.py hex 128 1 stiffness 8 27 7 | cut -c1-140
done
timeout 200 python bench.py --only-config 3 --fused-zero 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_kernel_ms'], d['parity']['ok'])"
