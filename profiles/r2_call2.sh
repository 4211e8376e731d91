# round 2, call 2: where does the time go on short rules?  consumers skipping work (CQ_DEBUG=1) / skipping the scatter (2)
set -x
for dbg in 0 1 2; do
 echo "== CQ_DEBUG=$dbg"
 CQ_DEBUG=$dbg timeout 200 python profiles/run_configs.py hexq1 quadnitsche 2>/dev/null | python -c "
import sys,json
for x in sys.stdin:
    try: d=json.loads(x)
    except: continue
    print('  ', d['case'][:34].ljust(34), d['form'].ljust(10), '%.3f ms'%d['ms'], 'frac %.3f'%d['frac_hbm'])"
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:assemble_ -s 2 -c 1 -f -o gpurun_out/r2_hexq1_stiff8 python profiles/one_form.py hex 128 1 stiffness 8 8 3 > gpurun_out/r2_ncu_a.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:assemble_ -s 2 -c 1 -f -o gpurun_out/r2_hexq1_scalar8 python profiles/one_form.py hex 128 1 scalar 8 8 3 > gpurun_out/r2_ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
