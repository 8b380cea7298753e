import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def seq384():
    """4 frames of a 384 x 384 synthetic sequence (mtf_b200/synth.py) + ground-truth warps."""
    from mtf_b200 import synth
    frames, warps = synth.make_sequence(4, 384, 384, seed=1234, walk_seed=5678, sigma=1.0)
    return frames, warps
