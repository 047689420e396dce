mkdir -p gpurun_out
export NRCU_TUNE_SETTINGS='[{"NRCU_TRACE_TAPER":"99,99"},{"NRCU_TRACE_TAPER":"6,12"},{"NRCU_TRACE_TAPER":"4,8"},{"NRCU_TRACE_TAPER":"3,6"},{"NRCU_TRACE_TAPER":"2,4"},{"NRCU_TRACE_TAPER":"99,99","NRCU_TRACE_BLOCKS":"4"},{"NRCU_TRACE_TAPER":"99,99"}]'
timeout 900 python tools/tune_trace.py 128 > gpurun_out/tune_v18.log 2>&1
cat gpurun_out/tune_v18.log
