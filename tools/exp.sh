M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
python bench.py > gpurun_out/r2d_bench_final.json 2> gpurun_out/r2d_bench_final.err
KIDMP_GRAPHS=0 $B > gpurun_out/r2d_plain_bench.log 2>&1 && KIDMP_GRAPHS=0 ncu --metrics $M --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r2d_ncu_launches.csv $B > gpurun_out/r2d_ncu_bench.log 2>&1
ls -la gpurun_out/
