import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The per-pixel candidate lists of the primary stage are built from 16 samples per pixel on (the break-even of their
# one walk per pixel); the parity tests render at 1-8 spp, so they lower the threshold to run WITH the lists
# (test_candidate_lists_on_off_and_default_are_bit_identical covers the switch itself).
os.environ.setdefault("RT_B200_PIXEL_LISTS_MIN_SPP", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def rtlib():
    """librt_b200.so through ctypes; building it needs only nvcc (no GPU)."""
    from raytracer_go_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(lib.LIB_PATH), "librt_b200.so"])
    return lib.load()


@pytest.fixture(scope="session")
def gpu(rtlib):
    """GPU tests fail loudly (never skip, never fall back) when no sm_100 device is usable."""
    n = rtlib.rt_device_count()
    assert n >= 1, "no sm_100 device visible: GPU tests need a B200 (librt_b200 has no CPU path)"
    return n


@pytest.fixture(scope="session")
def random_scene():
    from raytracer_go_b200 import scenes
    return scenes.random_scene()
