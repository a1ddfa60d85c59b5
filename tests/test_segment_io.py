"""YAML segment format of eth_trajectory_generation/io.cpp: round trip, 1 ns time quantisation, rejection of malformed documents."""
import os
import subprocess

import numpy as np

from mrs_uav_trajectory_generation_b200 import segment_io as IO


def test_round_trip_is_exact_in_coefficients_and_1ns_in_times(tmp_path):
    rng = np.random.default_rng(0)
    coef = rng.standard_normal((7, 4, 10)) * 10.0 ** rng.integers(-8, 8, size=(7, 4, 10))
    times = rng.uniform(0.01, 5.0, size=7)
    fn = str(tmp_path / "traj.yaml")
    assert IO.segments_to_file(fn, coef, times)
    c2, t2 = IO.segments_from_file(fn)
    assert np.array_equal(c2, coef)  # repr() of a double round-trips
    assert np.array_equal(t2, np.floor(times * 1e9).astype(np.uint64) * 1e-9)  # uint64 nanoseconds (segment.h:67-76)
    assert np.abs(t2 - times).max() <= 1e-9


def test_document_shape():
    text = IO.segments_to_yaml(np.ones((1, 4, 10)), [1.5])
    assert text.startswith("segments:\n  - N: 10\n    D: 4\n    time: 1500000000  # [ns]\n    coefficients:\n      - [1.0, ")
    assert IO.segments_from_yaml("foo: 1") is None
    assert IO.segments_from_yaml("segments:\n  - N: 10\n    D: 4\n    time: 5\n    coefficients:\n      - [1.0]\n") is None
    assert IO.segments_from_yaml("segments:\n  - N: 1\n    D: 2\n    time: 5\n    coefficients:\n      - [1.0]\n") is None
    c, t = IO.segments_from_yaml("segments:\n  - N: 2\n    D: 1\n    time: 2000000000\n    coefficients:\n      - [1.0, -2.5e-3]\n")
    assert c.shape == (1, 1, 2) and c[0, 0, 1] == -2.5e-3 and t[0] == 2.0


def test_cpp_header_reads_and_writes_the_same_documents(emu_lib, tmp_path):
    """include/eth_trajectory_generation_b200_io.hpp (segmentsToFile / segmentsFromFile / trajectoryTo/FromFile, io.cpp:125-218):
    reads what segment_io.py wrote, and what it writes back is read by segment_io.py with identical coefficients and times."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "tests", "host_emu")
    exe = str(tmp_path / "test_io")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-pthread", "-o", exe, os.path.join(root, "tests", "cpp", "test_io.cpp"),
                           "-L" + libdir, "-ltg_emu", "-Wl,-rpath," + libdir])
    rng = np.random.default_rng(5)
    coef = rng.standard_normal((9, 4, 10)) * 10.0 ** rng.integers(-12, 9, size=(9, 4, 10))
    coef[3, 2, 4] = 0.0
    coef[5, 0, 0] = -0.0
    times = rng.uniform(0.01, 9.0, size=9)
    a, b = str(tmp_path / "a.yaml"), str(tmp_path / "b.yaml")
    assert IO.segments_to_file(a, coef, times)
    out = subprocess.run([exe, a, b], capture_output=True, text=True, check=True).stdout.split()
    assert out[1] == "9" and out[3] == "5" and out[5] == "1"
    c2, t2 = IO.segments_from_file(b)
    t1 = np.floor(times * 1e9).astype(np.uint64) * 1e-9
    assert np.array_equal(c2.view(np.uint64), coef.view(np.uint64))  # bit for bit, the sign of zero included
    assert np.array_equal(t2, t1)                                    # a second trip through uint64 nanoseconds changes nothing more
    acc = 0.0
    for t in t1:
        acc += t
    assert float(out[7]) == acc
    assert out[9] == "1"  # an N = 6, D = 3 document: read, written back, evaluated through the general-shape entry point
