set -x
# per-kernel times and DRAM traffic of the config-4 step (owner-computes path), one rep each
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"owner|assemble_staged|DeviceScan" -c 40 --csv --log-file gpurun_out/r2s2_config4_launches.csv python examples/vector_poisson.py -config4 128 > gpurun_out/r2s2_config4_ncu.json 2> gpurun_out/r2s2_config4_ncu.err; echo rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2s2_config4_launches.csv')) if len(r)>10]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[1:]:
    print(r[ix['ID']], r[ix['Kernel Name']][:60], r[ix['Metric Name']], r[ix['Metric Value']], r[ix['Metric Unit']])
PY
