#!/usr/bin/env python
"""Research tool: per-ray BVH work (node steps, leaves, primitive tests) by bounce on the bench scene,
from the CPU emulation of the device code.  Usage: python tools/ray_work_stats.py [leaf_max]"""
import ctypes as C, os, subprocess, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from nrenderer_b200.flatscene import FlatScene

def main():
    extra = sys.argv[1:]
    so = "/tmp/libray_work_stats.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DNRCU_HOST_EMU=1"] + extra +
                   [f"-I{REPO}/include", f"-I{REPO}/nrenderer_b200/csrc", os.path.join(REPO, "tools", "ray_work_stats.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    L.emu_create.restype = C.c_void_p; L.emu_create.argtypes = [C.c_void_p, C.c_int]
    L.emu_ray_work.restype = C.c_uint32
    L.emu_ray_work.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    L.emu_bvh_stats.argtypes = [C.c_void_p, C.c_void_p]
    fs = FlatScene.load(os.path.join(REPO, "tests", "golden", "bunny5k_cornel.nrsc"))
    fs.width, fs.height, fs.samples_per_pixel, fs.depth, fs.cam_aspect = 1920, 1080, 1, 20, 16 / 9
    view, keep = fs.c_view()
    h = L.emu_create(C.addressof(view), 2)
    st = np.zeros(8, np.int32); L.emu_bvh_stats(h, st.ctypes.data); print("bvh: binary", st[0], "wide", st[1], "levels", st[2], "max_leaf", st[3], "n_big", st[6])
    rows = np.arange(0, 1080, 12)   # every 12th row, full width, pixel order
    pix = (rows[:, None] * 1920 + np.arange(1920)[None, :]).ravel().astype(np.uint32)
    cap = len(pix) * 8
    out = np.zeros((cap, 4), np.int32)
    n = L.emu_ray_work(h, 0, pix.ctypes.data, len(pix), 0, out.ctypes.data, cap)
    out = out[:n]
    print("rays", n, "paths", len(pix), "rays/path", n / len(pix))
    CN, CP = 115.0, 45.0
    for d in range(0, 8):
        m = out[out[:, 0] == d]
        if not len(m): break
        cost = CN * m[:, 1] + CP * m[:, 3]
        # SIMD efficiency if 32 consecutive rays ran in lock-step to the longest one (no refill)
        k = len(m) // 32 * 32
        c32 = cost[:k].reshape(-1, 32)
        eff = c32.mean() / c32.max(1).mean() if k else 0
        print(f"bounce {d}: rays {len(m):7d} nodes {m[:,1].mean():5.2f} (p50 {np.median(m[:,1]):.0f} p90 {np.percentile(m[:,1],90):.0f} p99 {np.percentile(m[:,1],99):.0f} max {m[:,1].max()}) "
              f"leaves {m[:,2].mean():4.2f} prims {m[:,3].mean():5.2f} (p90 {np.percentile(m[:,3],90):.0f} max {m[:,3].max()}) cost/ray {cost.mean():6.1f} warp-lockstep eff {eff:.2f}")
    allc = CN * out[:, 1] + CP * out[:, 3]
    print(f"all: nodes {out[:,1].mean():.2f} leaves {out[:,2].mean():.2f} prims {out[:,3].mean():.2f} ideal lane-instr/ray {allc.mean():.0f} = {allc.mean()/32:.1f} warp-instr/ray at 100% SIMD")
    print(f"rays entering the BVH: {(out[:,1]>0).mean()*100:.1f}%"); frac_heavy = (out[:, 1] > 8).mean(); print(f"rays with > 8 node steps: {frac_heavy*100:.1f}% carrying {allc[out[:,1]>8].sum()/allc.sum()*100:.1f}% of the work")
    # stage 1: candidates of the wide list per ray, and how well the balanced pass 2 of k_big_balanced fills its rounds
    L.emu_stage1_masks.restype = C.c_uint32
    L.emu_stage1_masks.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    sm = np.zeros((cap, 2), np.uint32)
    n1 = L.emu_stage1_masks(h, 0, pix.ctypes.data, len(pix), 0, sm.ctypes.data, cap)
    sm = sm[:n1]
    nb = int(st[6])
    bits = ((sm[:, 1][:, None] >> np.arange(nb)[None, :]) & 1).astype(np.int32)
    cand = bits.sum(1)
    print(f"stage 1: {nb} wide primitives; candidates per ray mean {cand.mean():.2f} p50 {np.median(cand):.0f} p90 {np.percentile(cand, 90):.0f} max {cand.max()}")
    print("         candidate rate per wide primitive:", " ".join(f"{x:.2f}" for x in bits.mean(0)))
    for d in (0, 1, 2, 5):
        m = cand[sm[:, 0] == d]
        k = len(m) // 32 * 32
        if not k: continue
        pairs = m[:k].reshape(-1, 32).sum(1)
        rounds = np.ceil(pairs / 32)
        print(f"         bounce {d}: pairs per warp {pairs.mean():6.1f}, rounds {rounds.mean():.2f}, lanes filled {pairs.sum() / (32 * rounds.sum()) * 100:.1f} % "
              f"(per-lane loop instead: {m[:k].reshape(-1, 32).max(1).mean():.2f} rounds at {m[:k].mean() / m[:k].reshape(-1, 32).max(1).mean() * 100:.0f} %)")

if __name__ == "__main__":
    main()
