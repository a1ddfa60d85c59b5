import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib as O

    O.build_oracle(ref=True)
    O.set_math_mode(O.MATH_DET)
    return O


@pytest.fixture(scope="session")
def emu_lib():
    """Host-emulation build of the product's device code (tests/host_emu): same C ABI, CPU threads."""
    import subprocess

    from mrs_uav_trajectory_generation_b200 import Library

    d = os.path.join(ROOT, "tests", "host_emu")
    subprocess.check_call(["make", "-s", "-C", d, "all"])
    return Library(os.path.join(d, "libtg_emu.so"))


@pytest.fixture(scope="session")
def emu_ctx(emu_lib):
    from mrs_uav_trajectory_generation_b200 import Context

    return Context(emu_lib, 0)


@pytest.fixture(scope="session")
def gpu_ctx():
    """The product: libtg_b200.so on cuda:0.  No fallback: the fixture fails if the library or the GPU is missing."""
    from mrs_uav_trajectory_generation_b200 import Context, Library

    return Context(Library(), 0)
