set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_owner.py tests/test_gpu_fused_zero.py -q -m gpu -x 2>&1 | tail -3
for lib in customquad_b200/libcq_b200.so build/variants/libcq_cb3.so build/variants/libcq_cbold.so; do echo "== $lib"
for r in "200 300" "100 120" "40 60" "8 27"; do set -- $r; CQ_LIB=$lib python profiles/one_form.py hex 96 1 stiffness $1 $2 5 | cut -c1-140; done
CQ_LIB=$lib python profiles/one_form.py hex 96 1 nitsche 40 60 5 | cut -c1-140
CQ_LIB=$lib python profiles/one_form.py hex 96 1 load 100 120 5 | cut -c1-140
CQ_LIB=$lib python profiles/one_form.py quad 1024 2 stiffness 200 300 5 | cut -c1-140
CQ_LIB=$lib timeout 300 python bench.py --only-config 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_kernel_ms'], d['parity']['ok'])"
done
