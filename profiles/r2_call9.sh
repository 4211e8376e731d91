set -x
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:assemble_ --csv --log-file gpurun_out/r2_c5_traffic_dyn.csv python profiles/config5_traffic.py 256 1000000000000 100000000 25000000 6000000 > gpurun_out/r2_c5_traffic_dyn.log 2>&1
python - <<'PY'
import csv
lines=open('gpurun_out/r2_c5_traffic_dyn.csv').read().splitlines()
i=[k for k,l in enumerate(lines) if l.startswith('"ID"')][0]
by={}
for r in csv.DictReader(lines[i:]):
    by.setdefault(r['ID'],{})[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
for k,v in by.items():
    print('dynamic', k, 'ms %.2f'%(v['gpu__time_duration.sum']/1e6), 'read GB %.1f'%(v['dram__bytes_read.sum']/1e9), 'write GB %.1f'%(v['dram__bytes_write.sum']/1e9), 'L2 hit %.1f'%v['lts__t_sector_hit_rate.pct'])
PY
timeout 300 python profiles/config5_single.py 256 2>&1 | grep -v Warning
for dyn in 0 1; do
  CQ_DYNAMIC=$dyn CQ_LIB=build/variants/libcq_prof.so timeout 120 python bench.py --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench dyn=$dyn', round(d['ms_per_step'],3), [(round(k['ms'],3),round(k['frac'],3)) for k in d['roofline']['per_kernel']])"
  echo "== hexq1 dyn=$dyn"
  CQ_DYNAMIC=$dyn CQ_LIB=build/variants/libcq_prof.so timeout 200 python profiles/run_configs.py hexq1 quadnitsche 2>/dev/null | python -c "
import sys,json
for x in sys.stdin:
    try: d=json.loads(x)
    except: continue
    print('  ', d['case'][:34].ljust(34), d['form'].ljust(10), '%.3f ms'%d['ms'], 'frac %.3f'%d['frac_hbm'])"
done
