import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import load_workload, DEFAULT_WORKLOAD
from nrenderer_b200 import Context
fs, mode, comp, bpr = load_workload(DEFAULT_WORKLOAD, 128)
ctx = Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
w, h = fs.width, fs.height
accum = torch.zeros(h, w, 4, device="cuda"); rgba = torch.empty(h, w, 4, device="cuda"); host = torch.empty(h, w, 4, pin_memory=True)
def t(f, n=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("upload ms", t(lambda: ctx.upload(fs, mode)))
print("render stats=True ms", t(lambda: (accum.zero_(), ctx.render_accumulate(accum.data_ptr(), s0=0, s1=128, want_stats=True))))
print("render stats=False ms", t(lambda: (accum.zero_(), ctx.render_accumulate(accum.data_ptr(), s0=0, s1=128, want_stats=False))))
print("resolve+d2h ms", t(lambda: (ctx.resolve(accum.data_ptr(), rgba.data_ptr()), host.copy_(rgba, non_blocking=True))))
def e2e():
    ctx.upload(fs, mode); accum.zero_(); ctx.render_accumulate(accum.data_ptr(), s0=0, s1=128, want_stats=False)
    ctx.resolve(accum.data_ptr(), rgba.data_ptr()); host.copy_(rgba, non_blocking=True); torch.cuda.synchronize()
print("e2e ms", t(e2e))
