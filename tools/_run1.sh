mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export LD_LIBRARY_PATH=$PWD/oracle/_ref:$PWD/nrenderer_b200:$LD_LIBRARY_PATH
for nd in 1 2; do
NRCU_DEVICES=$nd oracle/_ref/nr_headless --flat tests/golden/bunny5k_cornel.nrsc --w 1920 --h 1080 --aspect 1.7777778 --depth 20 --spp 256 --plugin nrenderer_b200/plugin/libNRCudaAccPathTracer.so --component CudaAccPathTracer --out gpurun_out/frame_nd$nd.f32 --repeat 2 2>&1 | tail -4
done
python - <<'PY'
import numpy as np
a=np.fromfile('gpurun_out/frame_nd1.f32',np.float32); b=np.fromfile('gpurun_out/frame_nd2.f32',np.float32)
print(a.shape, b.shape, float(np.abs(a-b).max()), float(a.mean()), float(b.mean()))
PY
rm -f gpurun_out/frame_nd*.f32
