import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
tests_dir = os.path.dirname(os.path.abspath(__file__))
if tests_dir not in sys.path:
    sys.path.insert(0, tests_dir)  # tests/patterns.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def uv_default():
    """The deterministic UV assignment fixed by SURVEY.md section 8d config 1: with the two shipped 16x16
    textures the atlas is 32x16, so the tiles are [0,0]-[0.5,1] and [0.5,0]-[1,1]
    (/root/reference/src/engine/rendering/texture_atlas.rs:68-96).  Row 0 (air) is unused."""
    import numpy as np
    return np.array([[0, 0, 0, 0], [0.0, 0.0, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]], dtype=np.float32)
