mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
export NRCU_TUNE_SETTINGS='[{}]'
for v in prev new prev new; do
  cp build/variants/libnrcuda_$v.so nrenderer_b200/libnrcuda.so
  echo "== $v" >> gpurun_out/tune_v17.log
  timeout 900 python tools/tune_trace.py 128 >> gpurun_out/tune_v17.log 2>&1
done
cat gpurun_out/tune_v17.log
