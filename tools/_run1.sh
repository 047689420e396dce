set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1b_bench_plain.json 2>gpurun_out/r1b_bench_plain.err; echo "plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r1b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1b_launches_bench.log 2>&1
tail -1 gpurun_out/r1b_launches_bench.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_big|k_trace2|k_shade' -s 3 -c 3 -f -o gpurun_out/r1b_hot_kernels python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r1b_hot_kernels.log 2>&1
tail -1 gpurun_out/r1b_hot_kernels.log | cut -c1-200
