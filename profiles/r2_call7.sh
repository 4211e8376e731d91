set -x
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "pattern_builders or plane_tile" 2>&1 | tail -3
timeout 600 python profiles/config5_single.py 256 2>&1 | grep -v Warning
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -k regex:assemble_ -s 1 -c 2 --csv --log-file gpurun_out/r2_c5_ncu.csv python profiles/config5_single.py 256 > gpurun_out/r2_c5_ncu.log 2>&1; cat gpurun_out/r2_c5_ncu.csv | tail -14
