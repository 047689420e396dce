set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_v4.log
tail -5 gpurun_out/pytest_v4.log
export NRCU_TUNE_SETTINGS='[{"NRCU_TRACE_VARIANT":"2","NRCU_TRACE_REFILL":"8"},{"NRCU_TRACE_VARIANT":"2","NRCU_TRACE_REFILL":"16"},{"NRCU_TRACE_VARIANT":"2","NRCU_TRACE_REFILL":"24"},{"NRCU_TRACE_VARIANT":"3","NRCU_TRACE_REFILL":"8"},{"NRCU_TRACE_VARIANT":"3","NRCU_TRACE_REFILL":"16"},{"NRCU_TRACE_VARIANT":"3","NRCU_TRACE_REFILL":"24"},{"NRCU_TRACE_VARIANT":"3","NRCU_TRACE_REFILL":"16","NRCU_TRACE_WNODE":"2","NRCU_TRACE_WPRIM":"1"}]'
timeout 900 python tools/tune_trace.py 64 > gpurun_out/tune_v4.log 2>&1
cat gpurun_out/tune_v4.log
