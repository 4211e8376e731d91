set -x
# session 3, call 6: the default policy on the bench workload; points in flight in the vector / scalar kernels
timeout 300 python bench.py --no-e2e --no-cpu --no-extra --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_form_ms'], d['config']['zeroing'][:40])"
timeout 300 python bench.py --no-e2e --no-cpu --no-extra --steps 5 --no-fused-zero 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_form_ms'], d['config']['zeroing'][:40])"
for lib in vs2 vs4; do echo "== $lib"
for c in "hex 96 1 load 100 120" "hex 96 1 load 8 27" "hex 96 1 load 8 8" "quad 1024 2 load 16 64" "quad 1024 1 load_coef 16 64" "hex 96 1 nitsche_rhs 4 16" "quad 1024 1 scalar 16 64" "hex 96 1 scalar 8 27"; do
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py $c 5 | cut -c1-140
done
done
CQ_LIB=build/variants/libcq_vs4.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2
