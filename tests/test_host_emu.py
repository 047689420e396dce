"""The device code (nrenderer_b200/csrc/*.cuh bodies) run on the CPU by tests/host_emu and checked
against the oracle: scene flattening, the BVH build + traversal, RayCast, the path tracer."""
import numpy as np
import pytest

import pyemu
from conftest import env_texture, glassify, load_scene, microfacet, random_rays
from oracle import pyoracle as po


@pytest.mark.parametrize("name,mode", [("ray_cast_cornel", 0), ("path_tracing_cornel", 1), ("bunny5k_cornel", 1), ("bunny5k_cornel", 2),
                                       ("pt_glass", 2), ("env_map_spheres", 2), ("bunny200_cornel", 2)])
def test_flattening_bounds_and_camera(name, mode):
    fs = load_scene(name, cam_aspect=1.5, cam_fov=170.0 if mode == 0 else 35.0)
    e, o = pyemu.EmuScene(fs, mode), po.OracleScene(fs, mode)
    ke, de, me, be = e.primitives()
    ko, do, mo = o.primitives()
    assert np.array_equal(ke, ko) and np.array_equal(me, mo)
    assert np.array_equal(be.view(np.uint32), o.bounds().view(np.uint32))
    if mode != 0:
        assert np.array_equal(de.view(np.uint32), do.view(np.uint32))
    ce, le = e.camera()
    co, lo = o.camera()
    assert np.array_equal(ce.view(np.uint32), co.view(np.uint32)) and le == lo


def test_bvh_build_invariants():
    fs = load_scene("bunny5k_cornel")
    st = pyemu.EmuScene(fs, 2).bvh_stats()
    assert st["leaf_slots"] == 4984 and st["max_leaf"] <= 4
    assert st["binary_nodes"] % 2 == 1 and st["wide_nodes"] < st["binary_nodes"]
    # every primitive reachable: a ray aimed at each primitive's centroid from just in front of it hits something at t>0
    tiny = load_scene("env_map_spheres")
    st = pyemu.EmuScene(tiny, 2).bvh_stats()
    # two spheres, each a sizeable part of the scene box: both go to the wide list and the tree is empty
    assert st["wide_nodes"] == 0 and st["n_big"] == 2 and st["root_ref"] == 0x7fffffff


@pytest.mark.parametrize("name,mode,n", [("path_tracing_cornel", 1, 60000), ("path_tracing_cornel", 2, 60000), ("bunny5k_cornel", 2, 60000),
                                         ("bunny5k_cornel", 1, 20000), ("bunny200_cornel", 2, 60000), ("env_map_spheres", 2, 20000),
                                         ("ray_cast_cornel", 0, 40000)])
def test_traversal_ids_bit_exact(name, mode, n):
    fs = load_scene(name)
    e, o = pyemu.EmuScene(fs, mode), po.OracleScene(fs, mode)
    rays = random_rays(n, seed=3 + mode)
    pe, te = e.trace_batch(rays)
    pid, t, tie = o.trace_batch(rays)
    assert (pe != pid).sum() == 0
    assert np.array_equal(te.view(np.uint32), t.view(np.uint32))
    if mode != 0:   # BVH == in-order brute force over the same packed records
        pl, tl = e.trace_batch(rays, linear=True)
        if mode == 1:
            assert np.array_equal(pl, pe)


def test_raycast_bit_exact():
    for (w, h) in [(500, 500), (97, 31)]:
        fs = load_scene("ray_cast_cornel", width=w, height=h)
        a, b = pyemu.EmuScene(fs, 0).render_raycast(), po.OracleScene(fs, 0).render_raycast()
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def lens(fs):
    fs.cam_aperture = 0.02


CASES = [("path_tracing_cornel", 1, 32, 32, 16, 4, 0, None), ("path_tracing_cornel", 2, 32, 32, 16, 20, 0, None),
         ("bunny5k_cornel", 2, 24, 24, 4, 20, 0, None), ("pt_glass", 2, 32, 32, 16, 20, 0, None),
         ("pt_glass", 2, 32, 32, 16, 8, 0, glassify), ("pt_glass", 2, 32, 32, 16, 8, 1, glassify),
         ("pt_glass_conductors", 2, 32, 32, 16, 8, 0, microfacet), ("env_map_spheres", 2, 48, 48, 16, 8, 0, env_texture),
         ("env_map_spheres", 2, 48, 48, 16, 8, 1, env_texture), ("path_tracing_cornel", 2, 32, 32, 16, 4, 0, lens)]


@pytest.mark.parametrize("name,mode,w,h,spp,depth,glass,edit", CASES)
def test_path_tracer_bit_exact_with_same_rng(name, mode, w, h, spp, depth, glass, edit):
    fs = load_scene(name, width=w, height=h, samples_per_pixel=spp, depth=depth)
    if edit:
        edit(fs)
    ae, re_ = pyemu.EmuScene(fs, mode).render_pt_accum(seed=7, glass_mode=glass)
    ao, ro = po.OracleScene(fs, mode).render_pt_accum(seed=7, glass_mode=glass)
    assert re_ == ro
    assert np.array_equal(ae.view(np.uint32), ao.view(np.uint32))
    assert np.isfinite(ae).all()


def test_depth_zero_and_sparse_pixels():
    fs = load_scene("path_tracing_cornel", width=16, height=16, samples_per_pixel=4, depth=0)
    fs.ambient_constant = np.array([0.1, 0.2, 0.3], np.float32)
    a, rays = pyemu.EmuScene(fs, 2).render_pt_accum()
    assert rays == 0 and np.allclose(a[..., :3] / a[..., 3:4], [0.1, 0.2, 0.3])
    fs = load_scene("bunny5k_cornel", width=1920, height=1080, samples_per_pixel=2, depth=20, cam_aspect=16 / 9)
    px = np.random.default_rng(0).choice(1920 * 1080, 64, replace=False).astype(np.uint32)
    a, _ = pyemu.EmuScene(fs, 2).render_pt_accum(seed=2, pixels=px)
    b, _ = po.OracleScene(fs, 2).render_pt_accum(seed=2, pixels=px)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_rejects_broken_scenes():
    fs = load_scene("path_tracing_cornel")
    fs.plane_material[0] = 77
    with pytest.raises(ValueError, match="material"):
        pyemu.EmuScene(fs, 2)


def _slice_means(emu, n_slices, spp, **kw):
    """Image-mean radiance of independent sample slices (linear space)."""
    out = []
    for k in range(n_slices):
        a, _ = emu.render_pt_accum(seed=9, s0=k * spp, s1=(k + 1) * spp, **kw)
        out.append((a[..., :3] / a[..., 3:4]).mean())
    return np.array(out)


@pytest.mark.parametrize("name,mode,depth", [("path_tracing_cornel", 2, 6), ("path_tracing_cornel", 1, 3), ("bunny200_cornel", 2, 5)])
def test_next_event_estimation_keeps_the_expectation(name, mode, depth):
    """NRCU_FLAG_NEE is an extension (the reference only hits lights by chance): same mean, far lower variance."""
    fs = load_scene(name, width=20, height=20, samples_per_pixel=16 * 96, depth=depth)
    e = pyemu.EmuScene(fs, mode)
    plain = _slice_means(e, 16, 96)
    nee = _slice_means(e, 16, 96, flags=1)
    sem = np.sqrt(plain.var(ddof=1) / len(plain) + nee.var(ddof=1) / len(nee))
    print(f"{name} m{mode}: plain {plain.mean():.5f} +- {plain.std(ddof=1) / 4:.5f}, NEE {nee.mean():.5f} +- {nee.std(ddof=1) / 4:.5f}")
    assert abs(plain.mean() - nee.mean()) < 4.5 * sem + 1e-4
    assert nee.std(ddof=1) < 0.75 * plain.std(ddof=1)         # lower variance even for the image mean (per pixel far more)
    # flags = 0 is bit-identical to the estimator without the extension compiled in
    a0, r0 = e.render_pt_accum(seed=3, s0=0, s1=4)
    a1, r1 = e.render_pt_accum(seed=3, s0=0, s1=4, flags=0)
    assert np.array_equal(a0, a1) and r0 == r1
