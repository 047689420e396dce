set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v11.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_v11.log
tail -3 gpurun_out/pytest_v11.log
export NRCU_TUNE_SETTINGS='[{"NRCU_TRACE_VARIANT":"2"},{"NRCU_TRACE_VARIANT":"4"},{"NRCU_TRACE_VARIANT":"4","NRCU_TRACE_REFILL":"16"},{"NRCU_TRACE_VARIANT":"4","NRCU_TRACE_REFILL":"4"},{"NRCU_TRACE_VARIANT":"3"}]'
timeout 900 python tools/tune_trace.py 128 > gpurun_out/tune_v11.log 2>&1
cat gpurun_out/tune_v11.log
