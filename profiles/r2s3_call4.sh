set -x
# session 3, call 4: threshold of the quadrilateral long-rule shape; full GPU test suite on the in-tree library
for lib in customquad_b200/libcq_b200.so build/variants/libcq_long32.so; do echo "== $lib"
CQ_LIB=$lib timeout 300 python bench.py --no-e2e --no-cpu --no-extra --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], [(k['kernel'], round(k['ms'],4), round(k['frac'],3)) for k in d['roofline']['per_kernel']])"
done
for lib in customquad_b200/libcq_b200.so build/variants/libcq_long48.so build/variants/libcq_long32.so; do echo "== $lib"
for r in "40 80" "50 90"; do set -- $r
CQ_LIB=$lib python profiles/one_form.py quad 1024 2 stiffness $1 $2 5 | cut -c1-140
CQ_LIB=$lib python profiles/one_form.py quad 1024 1 stiffness $1 $2 5 | cut -c1-140
done
done
python profiles/one_form.py hex 96 1 nitsche 4 16 5 | cut -c1-140
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
