mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_long_|k_enc_' --csv --log-file gpurun_out/long_launches.csv python scripts/prof_large.py --iters 1 > gpurun_out/ncu_ll.log 2>&1
tail -2 gpurun_out/ncu_ll.log
