"""Flat-scene plumbing: .nrsc round trips in Python and through the C++ harness / reference importers."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, REPO, load_scene
from nrenderer_b200.flatscene import FlatScene
from oracle import pyoracle as po


def test_python_round_trip(tmp_path):
    fs = load_scene("bunny5k_cornel", width=123, height=45, cam_aspect=2.5)
    fs.add_texture(np.random.default_rng(0).uniform(0, 1, (4, 8, 4)).astype(np.float32))
    p = tmp_path / "x.nrsc"
    fs.save(p)
    g = FlatScene.load(p)
    for name in FlatScene._ARRAYS:
        assert np.array_equal(np.asarray(getattr(fs, name)).reshape(-1), np.asarray(getattr(g, name)).reshape(-1)), name
    assert (g.width, g.height, g.cam_aspect) == (123, 45, 2.5)


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")
def test_harness_round_trip_through_the_reference_scene_type(tmp_path):
    """flat -> NRenderer::Scene (unflatten) -> flat (flatten) is the identity."""
    for name in ["bunny200_cornel", "pt_glass_conductors", "ray_cast_cornel"]:
        out = tmp_path / f"{name}.nrsc"
        env = dict(os.environ, LD_LIBRARY_PATH=po.REF_DIR)
        subprocess.run([os.path.join(po.REF_DIR, "nr_headless"), "--flat", os.path.join(GOLDEN, name + ".nrsc"), "--dump-flat", str(out)],
                       check=True, env=env)
        a, b = FlatScene.load(os.path.join(GOLDEN, name + ".nrsc")), FlatScene.load(out)
        for arr in FlatScene._ARRAYS:
            assert np.array_equal(np.asarray(getattr(a, arr)).reshape(-1), np.asarray(getattr(b, arr)).reshape(-1)), (name, arr)


@pytest.mark.skipif(not os.path.isdir("/root/reference/resource"), reason="reference resources not mounted")
def test_fixtures_are_what_the_reference_importers_produce(tmp_path):
    out = tmp_path / "s.nrsc"
    env = dict(os.environ, LD_LIBRARY_PATH=po.REF_DIR)
    subprocess.run([os.path.join(po.REF_DIR, "nr_headless"), "--scn", "/root/reference/resource/path_tracing_cornel.scn", "--dump-flat", str(out)],
                   check=True, env=env)
    a, b = load_scene("path_tracing_cornel"), FlatScene.load(out)
    for arr in FlatScene._ARRAYS:
        assert np.array_equal(np.asarray(getattr(a, arr)).reshape(-1), np.asarray(getattr(b, arr)).reshape(-1)), arr


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")
def test_cuda_plugins_register_through_the_reference_factory():
    """The adapters load next to the reference's own plugins and register under their names."""
    from nrenderer_b200 import build
    plugins = [build.plugin_path(m) for m in (0, 1, 2, 3)]
    if not all(os.path.exists(p) for p in plugins):
        pytest.skip("plugin adapters not built")
    cmd = [os.path.join(po.REF_DIR, "nr_headless"), "--list"]
    for p in plugins + [os.path.join(po.REF_DIR, n) for n in po.REF_PLUGINS.values()]:
        cmd += ["--plugin", p]
    env = dict(os.environ, LD_LIBRARY_PATH=po.REF_DIR)
    names = set(subprocess.run(cmd, capture_output=True, text=True, check=True, env=env).stdout.split())
    assert {"CudaRayCast", "CudaSimplePathTracer", "CudaAccPathTracer", "CudaMetropolisLightTransport", "RayCast", "SimplePathTracer", "AccPathTracer"} <= names


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")
def test_harness_image_writers(tmp_path):
    """--out by extension: raw RGBA fp32, .pfm (bottom-up RGB fp32) and .ppm / .png (8-bit) hold the same published frame."""
    env = dict(os.environ, LD_LIBRARY_PATH=po.REF_DIR)
    base = [os.path.join(po.REF_DIR, "nr_headless"), "--flat", os.path.join(GOLDEN, "ray_cast_cornel.nrsc"), "--w", "40", "--h", "30",
            "--plugin", os.path.join(po.REF_DIR, po.REF_PLUGINS["RayCast"]), "--component", "RayCast", "--out"]
    raw, pfm, ppm, png = tmp_path / "f.f32", tmp_path / "f.pfm", tmp_path / "f.ppm", tmp_path / "f.png"
    for o in (raw, pfm, ppm, png):
        subprocess.run(base + [str(o)], check=True, env=env, capture_output=True)
    img = np.fromfile(raw, np.float32).reshape(30, 40, 4)
    head, data = pfm.read_bytes().split(b"-1.0\n", 1)
    assert head == b"PF\n40 30\n"
    assert np.array_equal(np.frombuffer(data, np.float32).reshape(30, 40, 3)[::-1], img[..., :3])
    head, data = ppm.read_bytes().split(b"255\n", 1)
    assert head == b"P6\n40 30\n"
    want = (np.clip(img[..., :3], 0, 1) * 255 + 0.5).astype(np.uint8)
    assert np.array_equal(np.frombuffer(data, np.uint8).reshape(30, 40, 3), want)
    # PNG: signature, chunk CRCs, zlib stream (stored blocks + adler32) and the scanlines (filter byte 0)
    import struct, zlib
    blob = png.read_bytes()
    assert blob[:8] == b"\x89PNG\r\n\x1a\n"
    off, chunks = 8, []
    while off < len(blob):
        n, tag = struct.unpack(">I4s", blob[off:off + 8])
        body = blob[off + 8:off + 8 + n]
        assert struct.unpack(">I", blob[off + 8 + n:off + 12 + n])[0] == zlib.crc32(tag + body)
        chunks.append((tag, body)); off += 12 + n
    assert [t for t, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    assert struct.unpack(">IIBBBBB", chunks[0][1]) == (40, 30, 8, 2, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(30, 1 + 3 * 40)
    assert not rows[:, 0].any() and np.array_equal(rows[:, 1:].reshape(30, 40, 3), want)


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")
def test_component_manager_hosts_plugins_like_the_gui(tmp_path):
    """harness/ComponentManager.hpp (Linux stand-in for app/include/manager/ComponentManager.hpp): directory scan + dlopen,
    exec on a detached thread, IDLING -> READY -> RUNNING -> FINISH; the frame equals the direct exec() frame."""
    fs = load_scene("ray_cast_cornel", width=60, height=40)
    direct, d_info = po.run_reference(fs, "RayCast")
    managed, info = po.run_reference(fs, "RayCast", plugin_dirs=[po.REF_DIR], manager=True, repeat=2, warmup=1)
    assert "ComponentManager" in info["via"] and "RenderComponent::exec" in d_info["via"]
    assert info["repeat"] == 2 and info["warmup"] == 1 and info["screen_updates"] == 2 and info["errors"] == 0
    assert np.array_equal(direct, managed)
    # an unknown component is refused before a thread is started
    env = dict(os.environ, LD_LIBRARY_PATH=po.REF_DIR)
    r = subprocess.run([os.path.join(po.REF_DIR, "nr_headless"), "--flat", os.path.join(GOLDEN, "ray_cast_cornel.nrsc"), "--plugin-dir", po.REF_DIR,
                        "--manager", "--component", "NoSuchRenderer"], capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "not registered" in r.stderr


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")
def test_texture_files_go_through_the_reference_image_loader(tmp_path):
    """--texture decodes with the reference's ImageLoader (stb, channel / 255.f, 4 channels: ImageLoader.cpp:8-19) and
    --env-map sets Ambient::Type::ENVIROMENT_MAP + the texture handle like SceneBuilder.cpp:89-98."""
    import struct, zlib
    rng = np.random.default_rng(7)
    w, h = 20, 10
    tex8 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    raw = b"".join(b"\x00" + tex8[y].tobytes() for y in range(h))

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body))
    png = tmp_path / "env.png"
    png.write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))
    out = tmp_path / "t.nrsc"
    env = dict(os.environ, LD_LIBRARY_PATH=po.REF_DIR)
    subprocess.run([os.path.join(po.REF_DIR, "nr_headless"), "--flat", os.path.join(GOLDEN, "env_map_spheres.nrsc"), "--texture", str(png), "--env-map", "0",
                    "--dump-flat", str(out)], check=True, env=env)
    fs = FlatScene.load(out)
    assert fs.ambient_type == 1 and fs.ambient_environment_map == 0
    assert (int(fs.texture_width[0]), int(fs.texture_height[0])) == (w, h)
    tex = fs.texture_rgba.reshape(h, w, 4)
    assert np.array_equal(tex[..., :3], tex8.astype(np.float32) / np.float32(255.0)) and (tex[..., 3] == 1).all()
