set -x
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:assemble_ --csv --log-file gpurun_out/r2_c5_traffic.csv python profiles/config5_traffic.py 256 1000000000000 100000000 25000000 6000000 1500000 400000 > gpurun_out/r2_c5_traffic.log 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader(open('gpurun_out/r2_c5_traffic.csv')))
by={}
for r in rows:
    by.setdefault(r['ID'],{})[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
for k,v in by.items():
    print(k, 'ms %.2f'%(v['gpu__time_duration.sum']/1e6), 'read GB %.1f'%(v['dram__bytes_read.sum']/1e9), 'write GB %.1f'%(v['dram__bytes_write.sum']/1e9), 'L2 hit %.1f'%v['lts__t_sector_hit_rate.pct'])
PY
grep tile gpurun_out/r2_c5_traffic.log
timeout 300 python -m pytest tests/test_gpu_cutfem_poisson.py -q -x 2>&1 | tail -4
