mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
NRCU_TRACE_UPLOAD=1 python tools/_diag_e2e.py 1024 > gpurun_out/diag_e2e4.log 2>&1; grep -v "nrcu upload\]  \|host_prepare" gpurun_out/diag_e2e4.log | head -40
