"""The drop-in itself on the GPU: the REGISTER_RENDERER plugins (libNRCuda*.so) hosted by nr_headless exactly like the
reference's GUI hosts its components —

    Scene -> ComponentFactory::createComponent<RenderComponent>("Render", name) -> exec(onStart, onFinish, scene)
          -> Adapter::render(SharedScene) -> nrcu_* (C ABI) -> getServer().screen.set(...)

(reference code/server/component/RenderComponent.cpp:5-9, components/ray_cast/src/Adapter.cpp:11-34,
app/include/manager/ComponentManager.hpp:41-64).  Every other GPU test enters below the adapter through ctypes; these
enter where a user of the reference does.  Needs oracle/_ref (libNRServer.so + nr_headless, built by
__graft_entry__.build() where /root/reference is mounted; the binaries travel to the GPU box) — nothing here reads
/root/reference at run time.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, glassify, load_scene
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

MODE_PLUGIN = {0: "CudaRayCast", 1: "CudaSimplePathTracer", 2: "CudaAccPathTracer"}


def plugin(mode):
    from nrenderer_b200 import build
    p = build.plugin_path(mode)
    if not (po.ref_available() and os.path.exists(p)):
        pytest.fail("oracle/_ref or the plugin adapters were not built: run `python __graft_entry__.py` where /root/reference is mounted "
                    "(the drop-in cannot be tested without its host)")
    return p


def run_plugin(fs, mode, **kw):
    return po.run_reference(fs, MODE_PLUGIN[mode], extra_plugins=[plugin(mode)], timeout=900, **kw)


@pytest.fixture(scope="module")
def ctx():
    from nrenderer_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def test_raycast_plugin_frame_matches_the_reference_frame():
    """cfg1 through the plugin: 0 of 250 000 pixels beyond 1e-4 relative of the reference's own RayCast frame."""
    fs = load_scene("ray_cast_cornel")
    img, info = run_plugin(fs, 0)
    ref = np.load(os.path.join(GOLDEN, "ray_cast_cornel_500_ref.npz"))["rgb"]
    assert img.shape == (500, 500, 4) and (img[..., 3] == 1).all()
    rel = np.abs(img[..., :3] - ref) / np.maximum(np.abs(ref), 1e-6)
    print(f"CudaRayCast via {info['via']}: {info['seconds']:.4f} s, {(rel > 1e-4).any(-1).sum()} pixels beyond 1e-4; {info['last_log']}")
    assert (rel > 1e-4).sum() == 0
    assert info["errors"] == 0 and info["component"] == "CudaRayCast"


@pytest.mark.parametrize("name,mode,w,h,spp,depth", [("bunny5k_cornel", 2, 160, 90, 16, 20), ("path_tracing_cornel", 1, 96, 96, 24, 4),
                                                     ("pt_glass", 2, 128, 72, 16, 8)])
def test_path_tracer_plugin_frame_equals_the_c_abi_frame(ctx, name, mode, w, h, spp, depth):
    """What Screen holds after exec() is bit-for-bit what nrcu_render returns for the same flattened Scene and seed
    (Screen::set only clamps to [0,1], Screen.cpp:54-66)."""
    fs = load_scene(name, width=w, height=h, samples_per_pixel=spp, depth=depth, cam_aspect=w / h)
    img, info = run_plugin(fs, mode, env={"NRCU_SEED": "0"})
    ctx.upload(fs, mode)
    want, st = ctx.render(seed=0)
    assert info["errors"] == 0, info["last_error"]
    assert img.shape == want.shape
    assert np.array_equal(img.view(np.uint32), np.clip(want, 0.0, 1.0).view(np.uint32))
    assert "Mpath-samples/s" in info["last_log"]


def test_plugin_through_component_manager_detached_thread():
    """The GUI's route: ComponentManager::init scans a directory, exec runs the component on a detached thread, the host
    polls getState() and Screen::isUpdated() (nrenderer_b200/harness/ComponentManager.hpp mirrors
    app/include/manager/ComponentManager.hpp:15-70 without Win32)."""
    fs = load_scene("bunny200_cornel", width=96, height=64, samples_per_pixel=8, depth=6)
    direct, _ = run_plugin(fs, 2, env={"NRCU_SEED": "3"})
    img, info = po.run_reference(fs, "CudaAccPathTracer", plugin_dirs=[os.path.dirname(plugin(2))], manager=True, env={"NRCU_SEED": "3"}, timeout=900)
    assert "ComponentManager" in info["via"] and info["errors"] == 0
    assert info["screen_updates"] >= 1
    assert np.array_equal(img.view(np.uint32), direct.view(np.uint32))


def test_progressive_plugin_publishes_intermediate_frames_and_the_same_final_frame():
    fs = load_scene("path_tracing_cornel", width=64, height=48, samples_per_pixel=24, depth=6)
    one, _ = run_plugin(fs, 2, env={"NRCU_SEED": "2"})
    img, info = po.run_reference(fs, "CudaAccPathTracer", extra_plugins=[plugin(2)], manager=True,
                                 env={"NRCU_SEED": "2", "NRCU_PROGRESSIVE": "8"}, timeout=900)
    assert info["errors"] == 0
    np.testing.assert_allclose(img, one, rtol=2e-6, atol=1e-7)     # same samples; fp32 sums grouped per update
    print(f"progressive: {info['screen_updates']} screen updates seen by the polling host")
    assert info["screen_updates"] >= 1


def test_broken_scene_gives_a_black_frame_and_an_error_log():
    """Error convention of SURVEY 8(b): never throw out of render(); log error(...) and publish a black frame."""
    fs = load_scene("path_tracing_cornel", width=40, height=30, samples_per_pixel=2, depth=3)
    fs.node_entity[0] = 999                      # a node that points at a sphere that does not exist
    img, info = run_plugin(fs, 2)
    assert img.shape == (30, 40, 4)
    assert (img[..., :3] == 0).all() and (img[..., 3] == 1).all()
    assert info["errors"] >= 1 and "NRCuda" in info["last_error"]


def test_two_devices_through_the_plugin():
    import nrenderer_b200 as nr
    if nr.device_count() < 2:
        pytest.skip("needs two GPUs")
    fs = load_scene("bunny200_cornel", width=96, height=64, samples_per_pixel=12, depth=8)
    one, _ = run_plugin(fs, 2, env={"NRCU_SEED": "5"})
    two, info = run_plugin(fs, 2, env={"NRCU_SEED": "5", "NRCU_DEVICES": "2"})
    assert info["errors"] == 0 and "2 GPU(s)" in info["last_log"]
    np.testing.assert_allclose(two, one, rtol=3e-6, atol=1e-6)


def test_env_map_from_an_image_file_through_the_reference_loader(tmp_path):
    """the environment map as the reference's UI feeds it - an image file decoded by the reference's
    ImageLoader (stb -> RGBA / 255, app/src/utilities/ImageLoader.cpp:8-19), Ambient::Type::ENVIROMENT_MAP + texture
    handle (SceneBuilder.cpp:89-98) - gives the frame the C ABI renders from the same texels."""
    import subprocess
    import struct
    import zlib
    rng = np.random.default_rng(7)
    w, h = 48, 24
    tex8 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    raw = b"".join(b"\x00" + tex8[y].tobytes() for y in range(h))

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body))
    png = tmp_path / "env.png"
    png.write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))
    fs = load_scene("env_map_spheres", width=96, height=54, samples_per_pixel=8, depth=6, cam_aspect=16 / 9)
    scene, out, flat = tmp_path / "s.nrsc", tmp_path / "f.f32", tmp_path / "with_tex.nrsc"
    fs.save(scene)
    env = dict(os.environ, LD_LIBRARY_PATH=po.REF_DIR, NRCU_SEED="9")
    r = subprocess.run([os.path.join(po.REF_DIR, "nr_headless"), "--flat", str(scene), "--texture", str(png), "--env-map", "0", "--dump-flat", str(flat),
                        "--plugin", plugin(2), "--component", "CudaAccPathTracer", "--out", str(out)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    img = np.fromfile(out, np.float32).reshape(54, 96, 4)
    from nrenderer_b200 import api
    from nrenderer_b200.flatscene import FlatScene
    fs2 = FlatScene.load(flat)
    # ImageLoader semantics: 8-bit channels / 255, alpha 1 for a 3-channel file, row 0 = top row of the file
    tex = fs2.texture_rgba.reshape(h, w, 4)
    assert fs2.ambient_type == 1 and fs2.ambient_environment_map == 0
    assert np.array_equal(tex[..., :3], tex8.astype(np.float32) / np.float32(255.0)) and (tex[..., 3] == 1).all()
    c = api.Context(0)
    try:
        c.upload(fs2, 2)
        want, _ = c.render(seed=9)
    finally:
        c.close()
    assert np.array_equal(img.view(np.uint32), np.clip(want, 0, 1).view(np.uint32))
    assert img[..., :3].mean() > 0.05          # lit by the map (the reference renders this scene black)


def test_metropolis_plugin_registers_and_renders():
    """libNRCudaMetropolisLightTransport.so registers "CudaMetropolisLightTransport" (counterpart of the reference's
    MetropolisLightTransport, metropolis_light_transport/src/Adapter.cpp:36) and publishes a frame in the reference MLT's
    tone map; with NRCU_MLT_TONE=0 (sqrt) it is comparable with CudaSimplePathTracer's frame."""
    from nrenderer_b200 import build
    so = build.plugin_path(3)
    if not (po.ref_available() and os.path.exists(so)):
        pytest.fail("plugin adapters not built")
    fs = load_scene("path_tracing_cornel", width=48, height=40, samples_per_pixel=1024, depth=5, cam_aspect=1.2)
    img, info = po.run_reference(fs, "CudaMetropolisLightTransport", extra_plugins=[so], env={"NRCU_SEED": "4", "NRCU_MLT_TONE": "0"}, timeout=900)
    ref, _ = run_plugin(fs, 1, env={"NRCU_SEED": "4"})
    assert info["errors"] == 0 and img.shape == ref.shape
    a, b = img[..., :3].astype(np.float64) ** 2, ref[..., :3].astype(np.float64) ** 2     # undo the sqrt gamma (frames were clamped to [0,1] by Screen::set)
    dark = b.max(-1) < 0.9                                                                   # unclamped pixels
    rel = abs(a[dark].mean() - b[dark].mean()) / b[dark].mean()
    print(f"CudaMetropolisLightTransport via the plugin API: {info['seconds']:.3f} s; linear mean {a[dark].mean():.5f} vs CudaSimplePathTracer {b[dark].mean():.5f} ({rel * 100:.2f} %)")
    assert rel < 0.05
    toned, info2 = po.run_reference(fs, "CudaMetropolisLightTransport", extra_plugins=[so], env={"NRCU_SEED": "4"}, timeout=900)
    assert info2["errors"] == 0 and (toned[..., 3] == 1).all() and 0.05 < toned[..., :3].mean() < 0.95
