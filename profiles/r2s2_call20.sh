set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "hexahedra" 2>&1 | tail -2
timeout 600 python bench.py --only-config 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['per_kernel_ms'], d['roofline']['frac'], d['parity']['ok'])"
timeout 1200 python bench.py > gpurun_out/r2s2_bench_n1_final.json 2> gpurun_out/r2s2_bench_n1_final.err; echo rc=$?
tail -3 gpurun_out/r2s2_bench_n1_final.err
timeout 600 python bench.py --impl reference > gpurun_out/r2s2_bench_reference.json 2> gpurun_out/r2s2_bench_reference.err; echo rc=$?; cut -c1-400 gpurun_out/r2s2_bench_reference.json
