mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export LD_LIBRARY_PATH=$PWD/oracle/_ref:$PWD/nrenderer_b200:$LD_LIBRARY_PATH
NRCU_PROGRESSIVE=16 oracle/_ref/nr_headless --flat tests/golden/bunny5k_cornel.nrsc --w 640 --h 360 --aspect 1.7777778 --depth 20 --spp 64 --plugin nrenderer_b200/plugin/libNRCudaAccPathTracer.so --component CudaAccPathTracer --out gpurun_out/prog.ppm 2>&1 | tail -1 | cut -c1-300
ls -la gpurun_out/prog.ppm
