#!/bin/bash
# The ncu evidence of a round on one GPU (B200_PROFILING.md recipe): each capture only after the same command exited 0 without ncu.
#   TAG=r2b bash tools/profile_round.sh     -> gpurun_out/${TAG}_ncu_launches.csv, gpurun_out/${TAG}_prof.ncu-rep
TAG=${TAG:-r2b}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
B="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$B > gpurun_out/${TAG}_plain_bench.log 2>&1 || { echo "bench failed"; exit 1; }
ncu --metrics $M --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv $B > gpurun_out/${TAG}_ncu_bench.log 2>&1
P="python tools/profile_step.py --columns 1048576 --steps 2"
$P > gpurun_out/${TAG}_plain_step.log 2>&1 || { echo "profile_step failed"; exit 1; }
# (the step kernels of the SECOND step only: the report must stay under the 64 MiB that travel back)
ncu --set full --clock-control none --import-source on -k 'regex:^k_(classify|cell_|list_|n0_sweep|cells|carries|substeps|finish|diag_)' --launch-skip 15 -c 15 -o gpurun_out/${TAG}_prof -f $P > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_prof.ncu-rep gpurun_out/${TAG}_ncu_launches.csv
