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
