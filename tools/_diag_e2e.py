import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import load_workload, DEFAULT_WORKLOAD
from nrenderer_b200 import Context
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fs, mode, comp, bpr = load_workload(DEFAULT_WORKLOAD, spp)
ctx = Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
w, h = fs.width, fs.height
ctx.upload(fs, mode)
accum = torch.zeros(h, w, 4, device="cuda"); rgba = torch.empty(h, w, 4, device="cuda"); host = torch.empty(h, w, 4, pin_memory=True)
def sync(): torch.cuda.synchronize()
def timed(label, f):
    sync(); t0 = time.perf_counter(); r = f(); sync(); print(f"{label}: {(time.perf_counter()-t0)*1e3:.1f} ms", flush=True); return r
for it in range(2):
    timed("render stats=True ", lambda: (accum.zero_(), ctx.render_accumulate(accum.data_ptr(), s0=0, s1=spp, want_stats=True)))
for it in range(3):
    timed("upload            ", lambda: ctx.upload(fs, mode))
    timed("render stats=False", lambda: (accum.zero_(), ctx.render_accumulate(accum.data_ptr(), s0=0, s1=spp, want_stats=False)))
    timed("resolve+d2h       ", lambda: (ctx.resolve(accum.data_ptr(), rgba.data_ptr()), host.copy_(rgba, non_blocking=True)))
for it in range(2):
    timed("render stats=False (no upload)", lambda: (accum.zero_(), ctx.render_accumulate(accum.data_ptr(), s0=0, s1=spp, want_stats=False)))
for it in range(2):
    timed("render stats=True ", lambda: (accum.zero_(), ctx.render_accumulate(accum.data_ptr(), s0=0, s1=spp, want_stats=True)))
