# development: DRAM bytes of the Q2 mass kernel, memset + kernel (CQ_ZERO_FUSED=0) vs fused zeroing
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_write.sum
for z in 0 1; do
CQ_ZERO_FUSED=$z CQ_ZERO_AHEAD=${AHEAD:-8} timeout 200 ncu --metrics $M --clock-control none -k regex:assemble_moment_split -s 3 -c 1 --csv --log-file gpurun_out/dram_fz$z.csv python bench.py --no-e2e --no-cpu --steps 1 --forms q2_mass > gpurun_out/ncu_dram$z.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/dram_fz$z.csv")) if len(r)>5]
h=rows[0]
for r in rows[1:]: print("fused=$z", r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
done
