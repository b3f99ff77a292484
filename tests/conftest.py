import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ASSETS = os.path.join(ROOT, "tests", "assets")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box)")


@pytest.fixture(scope="session")
def std_env():
    from vk_gltf_renderer_b200 import hdr
    return hdr.load_hdr(os.path.join(ASSETS, "std_env.hdr"))


@pytest.fixture(scope="session")
def box_scene():
    from vk_gltf_renderer_b200 import scene
    return scene.load_gltf(os.path.join(ASSETS, "Box.glb"))


@pytest.fixture(scope="session")
def shader_ball_scene():
    from vk_gltf_renderer_b200 import scene
    return scene.load_gltf(os.path.join(ASSETS, "shader_ball.gltf"))


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


def rel_rmse(a, b):
    """sqrt(mean((a-b)^2)) / mean(b) over RGB."""
    a = np.asarray(a, np.float64)[..., :3]
    b = np.asarray(b, np.float64)[..., :3]
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.mean(b), 1e-30))
