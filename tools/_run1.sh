mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
export NRCU_TUNE_SETTINGS='[{"NRCU_WAVE_SKEW_KB":"0"},{"NRCU_WAVE_SKEW_KB":"1156"},{"NRCU_WAVE_SKEW_KB":"0","NRCU_BIG_BALANCED":"1"},{"NRCU_WAVE_SKEW_KB":"1156","NRCU_BIG_BALANCED":"1"},{"NRCU_WAVE_SKEW_KB":"68"}]'
timeout 900 python tools/tune_trace.py 128 > gpurun_out/tune_v16.log 2>&1
cat gpurun_out/tune_v16.log
