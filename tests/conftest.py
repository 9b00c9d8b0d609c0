import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: multi-second CPU test")


@pytest.fixture(scope="session")
def built():
    """Builds the in-tree native libraries once (nvcc cross-compiles without a GPU)."""
    from blitzen_b200 import build
    build.build_all()
    import oracle_lib
    oracle_lib.build()
    return True


@pytest.fixture(scope="session")
def tables(built):
    from blitzen_b200 import scene
    return scene.mesh_tables()


@pytest.fixture(scope="session")
def views(built):
    from blitzen_b200 import scene
    return scene.reference_views()


@pytest.fixture(scope="session")
def small_scene(built):
    """65 537 objects of the scaled stress mix (counter PRNG) in a 760-unit cube: several tiles plus a ragged tail."""
    from blitzen_b200 import scene
    return scene.stress_scene(n_stress=65_537 - 1001, multiplier=760.0, prng="counter", seed=7)


@pytest.fixture(scope="session")
def medium_scene(built):
    """1 048 577 objects: > 1000 tiles, exercises the look-back across many CTAs."""
    from blitzen_b200 import scene
    return scene.stress_scene(n_stress=1_048_577 - 1001, multiplier=1900.0, prng="counter", seed=11)


def view_at(position, yaw=0.0, pitch=0.0, z_far=650.0, width=1280, height=720):
    from blitzen_b200 import scene
    return scene.make_view(position, yaw, pitch, 70.0, width, height, 0.1, z_far)
