set -x
mkdir -p gpurun_out
export NRCU_TUNE_SETTINGS='[{"NRCU_FUSE_STAGE1":"0"}]'
for mb in 4 5 6; do
  cp build/variants/libnrcuda_mb$mb.so nrenderer_b200/libnrcuda.so
  echo "== minblocks $mb" >> gpurun_out/tune_v9.log
  timeout 900 python tools/tune_trace.py 64 >> gpurun_out/tune_v9.log 2>&1
done
cat gpurun_out/tune_v9.log
