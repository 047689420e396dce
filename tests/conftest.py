import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "host_emu"))
sys.path.insert(0, os.path.join(REPO, "tests"))
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_scene(name, **over):
    from nrenderer_b200.flatscene import FlatScene
    fs = FlatScene.load(os.path.join(GOLDEN, name + ".nrsc"))
    for k, v in over.items():
        setattr(fs, k, v)
    return fs


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def random_rays(n, seed=0):
    """Camera-like rays plus rays from random points inside the Cornell box in random directions."""
    rng = np.random.default_rng(seed)
    o = np.empty((n, 3), np.float32)
    d = np.empty((n, 3), np.float32)
    o[:, 0] = rng.uniform(-270, 270, n); o[:, 1] = rng.uniform(-270, 270, n); o[:, 2] = rng.uniform(760, 1300, n)
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    d[:] = v
    k = n // 4
    o[:k] = [0, 0, 10]
    d[:k, 0] = rng.uniform(-0.36, 0.36, k); d[:k, 1] = rng.uniform(-0.36, 0.36, k); d[:k, 2] = 1
    d[:k] /= np.linalg.norm(d[:k], axis=1, keepdims=True)
    # a few degenerate directions: axis aligned, zero components, the glass shader's (1,1,1)
    special = np.array([[1, 0, 0], [0, -1, 0], [0, 0, 1], [1, 1, 1], [0, 1, 1], [-1, 0, 1]], np.float32)
    m = min(len(special) * 50, n - k)
    d[k:k + m] = np.tile(special, (50, 1))[:m]
    return np.concatenate([o, d], 1).astype(np.float32)


def glassify(fs):
    fs.sphere_material[:] = fs.add_material(2, ior=1.5, absorbed=[1, 1, 1])


def microfacet(fs):
    fs.triangle_material[:] = 4 + 6
    fs.plane_material[5:] = 4 + 0


def env_texture(fs, w=64, h=32, seed=1):
    rng = np.random.default_rng(seed)
    tex = rng.uniform(0.05, 1.5, (h, w, 4)).astype(np.float32)
    fs.add_texture(tex)
    fs.ambient_type = 1
    fs.ambient_environment_map = 0


def lens_small(fs):
    fs.cam_aperture = 0.02            # lens radius 0.01 at the default focus distance 0.1


def lens_wide(fs):
    fs.cam_aperture = 6.0             # a visibly defocused view: lens radius 3, focused on the back of the box
    fs.cam_focus_distance = 900.0


def env_sun(fs, w=64, h=32):
    """Environment map with a small very bright region (a 'sun') over a dim sky: the case importance sampling exists for."""
    rng = np.random.default_rng(5)
    tex = rng.uniform(0.02, 0.2, (h, w, 4)).astype(np.float32)
    tex[5:8, 10:14, :3] = 60.0
    tex[20:22, 40:50, :3] = [8, 4, 1]
    fs.add_texture(tex)
    fs.ambient_type, fs.ambient_environment_map = 1, 0


def env_smooth(fs, w=64, h=32):
    y, x = np.linspace(0, 1, h)[:, None], np.linspace(0, 1, w)[None, :]
    tex = np.stack([0.5 + 0.5 * np.sin(6.28 * x) * np.ones_like(y), 0.3 + 0.7 * y * np.ones_like(x), 0.2 + 0.8 * (x * y), np.ones((h, w))], -1)
    fs.add_texture(tex.astype(np.float32))
    fs.ambient_type, fs.ambient_environment_map = 1, 0


def diffuse_spheres(fs):
    fs.sphere_material[:] = fs.add_material(0, diffuse_color=[0.7, 0.6, 0.5])
