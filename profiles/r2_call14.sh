set -x
for tile in 25165824 1000000000000; do
CQ_L2_TILE_BYTES=$tile timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none -k regex:assemble_staged -c 2 --csv --log-file gpurun_out/r2_c4_traffic_$tile.csv python examples/vector_poisson.py -config4 128 > /dev/null 2>&1
python - <<PY
import csv
lines=open('gpurun_out/r2_c4_traffic_$tile.csv').read().splitlines()
i=[k for k,l in enumerate(lines) if l.startswith('"ID"')][0]
by={}
for r in csv.DictReader(lines[i:]):
    by.setdefault(r['ID'],{})[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
for k,v in by.items():
    print('tile=$tile', k, 'ms %.2f'%(v['gpu__time_duration.sum']/1e6), 'read GB %.1f'%(v['dram__bytes_read.sum']/1e9), 'write GB %.1f'%(v['dram__bytes_write.sum']/1e9), 'L2 hit %.1f'%v['lts__t_sector_hit_rate.pct'], 'lsu %.1f'%v['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'], 'fp64 %.1f'%v['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'])
PY
done
