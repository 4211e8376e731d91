# development: fused-zeroing A/B (CQ_ZERO_FUSED=0: memset + kernel) over the look-ahead, then an ncu launch list
timeout 250 python -m pytest tests/test_gpu_fused_zero.py -q -m gpu -x > gpurun_out/fz_tests.log 2>&1; tail -2 gpurun_out/fz_tests.log
CQ_ZERO_FUSED=0 timeout 90 python bench.py --no-e2e --no-cpu > gpurun_out/bench_fz0.json 2> gpurun_out/bench_fz0.err
for a in ${AHEADS:-2 4 8}; do CQ_ZERO_AHEAD=$a timeout 90 python bench.py --no-e2e --no-cpu > gpurun_out/bench_fz1_a$a.json 2> gpurun_out/bench_fz1_a$a.err; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_fz*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), [round(v,3) for v in d["per_form_ms"].values()], [(round(k["ms"],3),round(k["frac"],3)) for k in d["roofline"]["per_kernel"]])
    except Exception as e: print(f, "ERR", e)
PY
CQ_ZERO_AHEAD=${NCU_AHEAD:-4} timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_fz.csv python bench.py --no-e2e --no-cpu --steps 1 > gpurun_out/ncu_fz.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_fz.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:90]].append(float(r[vi].replace(",","")))
for k,v in agg.items(): print(len(v), round(sum(v)/len(v)/1e3,1), "us", k)
PY
