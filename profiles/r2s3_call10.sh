set -x
# session 3, call 10: the wait for the zeroed ranges in the first moment warp instead of the expansion warps (Q2 stiffness with the zeroing warp)
run() { CQ_LIB=$1 timeout 300 python bench.py --no-e2e --no-cpu --no-extra --steps 5 $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_form_ms'])"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "device_resident" 2>&1 | tail -2
echo "== zbase default"; run build/variants/libcq_zbase.so ""
echo "== zbase fused all"; run build/variants/libcq_zbase.so --fused-zero
echo "== zmom default"; run build/variants/libcq_zmom.so ""
echo "== zmom fused all"; run build/variants/libcq_zmom.so --fused-zero
CQ_LIB=build/variants/libcq_zmom.so timeout 600 python -m pytest tests/test_gpu_fused_zero.py -q -m gpu -x 2>&1 | tail -2
