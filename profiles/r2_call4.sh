# round 2, call 4: FP64 peak, GPU tests, full bench at N=1 with the new legs
set -x
profiles/microbench/fp64_peak > gpurun_out/fp64_peak.json; cat gpurun_out/fp64_peak.json
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest4.log 2>&1; tail -5 gpurun_out/r2_pytest4.log
timeout 600 python bench.py > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; tail -5 gpurun_out/r2_bench4.err
timeout 300 python bench.py --impl reference > gpurun_out/r2_bench4_ref.json 2> gpurun_out/r2_bench4_ref.err; tail -2 gpurun_out/r2_bench4_ref.err
