"""Steady-state throughput of the regeneration kernels: the first NRCU_REGEN_MAX_ITERS iterations of a 1024-spp frame
(every slot alive), rays = iterations x slots."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nrenderer_b200 import Context
fs, mode, _ = bench.load_workload(bench.DEFAULT_WORKLOAD)
ctx = Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.upload(fs, mode)
acc = torch.zeros(fs.height, fs.width, 4, device="cuda")
for rep in range(3):
    acc.zero_()
    st = ctx.render_accumulate(acc.data_ptr(), seed=0, want_stats=True, scheduler=2, flags=4)
rays = st["iterations"] * st["max_queue"]
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("NRCU_")}, "iterations": st["iterations"], "slots": st["max_queue"], "ms": st["ms_total"],
                  "grays_per_s": rays / st["ms_total"] * 1e-6, "ms_trace": st["ms_trace"], "ms_stage2": st["ms_stage2"], "ms_shade": st["ms_shade"]}))
