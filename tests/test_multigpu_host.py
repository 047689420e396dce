"""N>1 host logic on CPU: sample slices + gloo reduce + resolve (world size 2 and 3), with the CPU emulation of
the kernels standing in for the GPU renderer.  The GPU run of the same code path is `bench.py --gpus N`."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, load_scene

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests", "host_emu"))

from nrenderer_b200 import multigpu  # noqa: E402


def test_sample_slices_partition_the_samples():
    for spp in (0, 1, 5, 16, 1024, 4096):
        for world in (1, 2, 3, 4, 8):
            sl = multigpu.all_slices(spp, world)
            assert sl[0][0] == 0 and sl[-1][1] == spp
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multigpu.sample_slice(16, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _resolve(accum, rgba):
    rgba[..., :3] = torch.sqrt(accum[..., :3] / accum[..., 3:4])
    rgba[..., 3] = 1.0


def _worker(rank, world, port, spp, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pyemu
    fs = load_scene("bunny200_cornel")
    fs.width, fs.height, fs.samples_per_pixel, fs.depth = 24, 16, spp, 6
    emu = pyemu.EmuScene(fs, 2)

    def render_slice(accum, s0, s1):
        a, rays = emu.render_pt_accum(seed=7, s0=s0, s1=s1)
        accum += torch.from_numpy(a)
        return {"rays": rays}

    accum = torch.zeros(fs.height, fs.width, 4)
    rgba = torch.zeros(fs.height, fs.width, 4) if rank == 0 else None
    st = multigpu.render_frame(render_slice, _resolve, accum, rgba, spp, rank, world)
    rays = torch.tensor([st["rays"] if st else 0], dtype=torch.int64)
    dist.all_reduce(rays)
    if rank == 0:
        np.savez(out_path, rgba=rgba.numpy(), accum=accum.numpy(), rays=rays.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,spp", [(2, 6), (3, 4), (4, 2)])
def test_slices_reduce_to_the_single_process_frame(tmp_path, world, spp):
    import pyemu
    out = str(tmp_path / "frame.npz")
    mp.spawn(_worker, args=(world, _free_port(), spp, out), nprocs=world, join=True)
    got = np.load(out)
    fs = load_scene("bunny200_cornel")
    fs.width, fs.height, fs.samples_per_pixel, fs.depth = 24, 16, spp, 6
    full, rays = pyemu.EmuScene(fs, 2).render_pt_accum(seed=7, s0=0, s1=spp)
    assert int(got["rays"][0]) == rays                         # the union of the slices is the same set of paths
    assert np.array_equal(got["accum"][..., 3], full[..., 3])   # every pixel got spp samples
    np.testing.assert_allclose(got["accum"][..., :3], full[..., :3], rtol=2e-6, atol=1e-6)   # fp32 summation order only
    ref = np.sqrt(full[..., :3] / full[..., 3:4])
    np.testing.assert_allclose(got["rgba"][..., :3], ref, rtol=2e-6, atol=1e-6)
    assert np.all(got["rgba"][..., 3] == 1.0)
