set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 1 --warmup 1 --spp 4 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_v4.log 2>&1
tail -2 gpurun_out/ncu_v4.log
