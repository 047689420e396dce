set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v7.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_v7.log
tail -5 gpurun_out/pytest_v7.log
export NRCU_TUNE_SETTINGS='[{"NRCU_FUSE_STAGE1":"0"},{"NRCU_FUSE_STAGE1":"1"}]'
timeout 900 python tools/tune_trace.py 64 > gpurun_out/tune_v7.log 2>&1
cat gpurun_out/tune_v7.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 1 --warmup 1 --spp 4 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_v7.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 2 -f -o gpurun_out/prof_shade_v7 python bench.py --steps 1 --warmup 1 --spp 4 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_shade_v7.log 2>&1
