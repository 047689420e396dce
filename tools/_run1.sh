mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s -k "full_resolution" 2>&1 | grep -v "^$" | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
