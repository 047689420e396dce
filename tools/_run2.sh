set -x
mkdir -p gpurun_out
export NRCU_TRACE_REFILL=16
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_trace3 -c 3 -f -o gpurun_out/prof_v3 python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_v3.log 2>&1
tail -3 gpurun_out/ncu_v3.log
