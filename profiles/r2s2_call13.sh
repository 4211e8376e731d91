set -x
timeout 600 python -m pytest tests/test_gpu_owner.py tests/test_gpu_parity.py -x -q -m gpu -k "owner or overwrite or reproducible or 3-2" 2>&1 | tail -3
for lib in customquad_b200/libcq_b200.so build/variants/libcq_ncw14.so build/variants/libcq_ncw10.so; do echo "== $lib"; CQ_LIB=$lib CQ_BS=3 CQ_ZERO=1 python profiles/one_form.py hex 48 2 stiffness 27 27 5; CQ_LIB=$lib CQ_BS=1 CQ_ZERO=0 python profiles/one_form.py hex 48 2 stiffness 27 27 5; done
timeout 600 python examples/vector_poisson.py -config4 128 > gpurun_out/r2s2_config4_owner5.json 2> gpurun_out/r2s2_config4_owner5.err; echo rc=$?; tail -3 gpurun_out/r2s2_config4_owner5.err; cat gpurun_out/r2s2_config4_owner5.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"owner|assemble_staged|DeviceScan" -c 60 --csv --log-file gpurun_out/r2s2_config4_launches2.csv python examples/vector_poisson.py -config4 128 > /dev/null 2> gpurun_out/r2s2_config4_ncu2.err; echo rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2s2_config4_launches2.csv')) if len(r)>10]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[1:]:
    print(r[ix['ID']], r[ix['Kernel Name']][:50], r[ix['Metric Name']], r[ix['Metric Value']], r[ix['Metric Unit']])
PY
