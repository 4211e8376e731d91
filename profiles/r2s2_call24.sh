set -x
for lib in build/variants/libcq_s2w10.so build/variants/libcq_s2w12.so build/variants/libcq_s3w10.so; do echo "== $lib"
for r in "100 120" "40 60" "200 300"; do set -- $r; CQ_LIB=$lib python profiles/one_form.py hex 96 1 stiffness $1 $2 5 | cut -c1-140; done
CQ_LIB=$lib timeout 300 python bench.py --only-config 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_kernel_ms'], d['parity']['ok'])"
done
