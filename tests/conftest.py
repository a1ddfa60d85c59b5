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


def _context(library, solve_kernels):
    """The library sends a solve launch to the thread-per-instance kernel when it has at least TG_THREAD_MIN_INST instances (default
    16384) and to the lane-parallel kernels otherwise; the knob is read when a context is created.  Every parity test runs both ways:
    "by-size" = the shipped dispatch (test batches are small: lane-parallel kernels, the large-batch tests reach the thread kernel),
    "thread" = thread-per-instance kernel for everything it can take, dense launches and work lists alike,
    "thread+octet-lists" (emulator only) = thread kernel for the dense launches, lane-parallel kernels over the work lists of the
    evaluation tail (what a large batch does once a list is shorter than TG_LIST_OCTET_BELOW)."""
    from mrs_uav_trajectory_generation_b200 import Context

    knobs = {"by-size": {}, "thread": {"TG_THREAD_MIN_INST": "0", "TG_LIST_OCTET_BELOW": "0"},
             "thread+octet-lists": {"TG_THREAD_MIN_INST": "0", "TG_LIST_OCTET_BELOW": "1000000000"}}[solve_kernels]
    names = ("TG_THREAD_MIN_INST", "TG_LIST_OCTET_BELOW")
    old = {k: os.environ.get(k) for k in names}
    for k in names:
        os.environ.pop(k, None)
    os.environ.update(knobs)
    try:
        ctx = Context(library, 0)
        ctx.solve_kernels = solve_kernels
        return ctx
    finally:
        for k in names:
            if old[k] is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = old[k]


@pytest.fixture(scope="session", params=["by-size", "thread", "thread+octet-lists"])
def emu_ctx(emu_lib, request):
    return _context(emu_lib, request.param)


@pytest.fixture(scope="session", params=["by-size", "thread"])
def gpu_ctx(request):
    """The product: libtg_b200.so on cuda:0.  No fallback: the fixture fails if the library or the GPU is missing."""
    from mrs_uav_trajectory_generation_b200 import Library

    return _context(Library(), request.param)
