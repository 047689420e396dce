# Captures the committed profiles (profiles/r2_*): run under gpurun, then tools/make_profile_summary.py r2 gpurun_out/f_launches.csv <raw csv of f_prof.ncu-rep> and tools/sass_hotspots.py
set -e
export NRCU_WAVE_MSLOTS=128
CMD="python bench.py --steps 1 --warmup 1 --spp 128 --no-cpu-baseline --no-e2e --no-other-workloads"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
$CMD > gpurun_out/f_plain.log 2>&1
ncu --metrics $M --clock-control none -c 640 --csv --log-file gpurun_out/f_launches.csv $CMD > gpurun_out/f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_shade_pool|k_big_balanced64|k_trace2|k_raygen" -s 0 -c 8 -f -o gpurun_out/f_prof $CMD > gpurun_out/f_ncu2.log 2>&1
tail -2 gpurun_out/f_ncu2.log
