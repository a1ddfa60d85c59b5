"""The C-ABI library loads and exports every symbol include/tg_b200.h declares (no compute calls: runs without a GPU),
and fails loudly -- never falls back -- when no CUDA device is usable."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for s in ("tg_ctx_create", "tg_optimize_batch", "tg_fetch_outputs", "tg_solve_linear_batch", "tg_time_alloc_batch", "tg_sample_batch",
              "tg_evaluate_batch", "tg_extrema_batch", "tg_max_magnitude_batch", "tg_objective_batch", "tg_scale_times_batch", "tg_sweep_costs"):
        assert s in syms


def test_cuda_library_exports_every_declared_symbol():
    from mrs_uav_trajectory_generation_b200 import build

    lib = C.CDLL(build.build())  # nvcc cross-compiles sm_100a without a GPU; loading needs libcudart only
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.tg_version.restype = C.c_char_p
    assert b"sm_100a" in lib.tg_version()


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mrs_uav_trajectory_generation_b200 import Context, Library, TgError

    with pytest.raises(TgError):
        Context(Library(), 0)


def test_package_does_not_touch_the_oracle():
    """Product code must never import, link or execute anything under oracle/ (tests, smoke() and bench's cpu_baseline only)."""
    pkg = os.path.join(ROOT, "mrs_uav_trajectory_generation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "oracle/" not in text.replace("the oracle/", ""), f
