"""YAML interchange of polynomial segments in the format of eth_trajectory_generation/io.cpp (data format):

    segments:
      - N: 10
        D: 4
        time: 1234567890  # [ns]
        coefficients:
          - [c0, c1, ..., c9]      one flow sequence per dimension, increasing powers
          - ...

Reference: io.cpp:27-31 (keys), 125-168 segmentsToFile, 169-218 segmentsFromFile, 46-122 the YAML::Node forms;
segment.h:67-76: the time is stored as uint64 nanoseconds, static_cast<uint64_t>(1e9 * t) on write and ns * 1e-9 on read, so a
round trip quantises segment times to 1 ns (the same loss as in the reference).  Host-side only: no GPU work here.
"""
import numpy as np
import yaml

K_SEGMENTS, K_N, K_D, K_TIME, K_COEFFS = "segments", "N", "D", "time", "coefficients"


def segments_to_yaml(coef, times):
    """coef [S][D][N] (increasing powers), times [S] seconds -> YAML text (segmentsToFile, io.cpp:125-168)."""
    coef = np.asarray(coef, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    if coef.ndim != 3 or coef.shape[0] != times.shape[0]:
        raise ValueError("coef must be [S][D][N] with one time per segment")
    lines = [K_SEGMENTS + ":"]
    for s in range(coef.shape[0]):
        lines.append("  - %s: %d" % (K_N, coef.shape[2]))
        lines.append("    %s: %d" % (K_D, coef.shape[1]))
        lines.append("    %s: %d  # [ns]" % (K_TIME, int(np.uint64(1.0e9 * times[s]))))  # getTimeNSec (segment.h:67-69)
        lines.append("    %s:" % K_COEFFS)
        for d in range(coef.shape[1]):
            lines.append("      - [" + ", ".join(repr(float(c)) for c in coef[s, d]) + "]")
    return "\n".join(lines) + "\n"


def segments_from_yaml(text):
    """YAML text -> (coef [S][D][N], times [S]); None when the document does not have the reference's shape
    (segmentsFromFile returns false, io.cpp:169-218)."""
    node = yaml.safe_load(text)
    if not isinstance(node, dict) or K_SEGMENTS not in node or not isinstance(node[K_SEGMENTS], list):
        return None
    coefs, times = [], []
    for seg in node[K_SEGMENTS]:
        if not isinstance(seg, dict) or any(k not in seg for k in (K_N, K_D, K_TIME, K_COEFFS)):
            return None
        n, d = int(seg[K_N]), int(seg[K_D])
        rows = seg[K_COEFFS]
        if not isinstance(rows, list) or len(rows) != d:
            return None  # "Coefficients and dimensions do not coincide."
        if any((not isinstance(r, list)) or len(r) != n for r in rows):
            return None  # "Number of coefficients does not match segment N."
        coefs.append(np.array(rows, dtype=np.float64))
        times.append(int(seg[K_TIME]) * 1.0e-9)  # setTimeNSec (segment.h:74-76)
    if coefs and any(c.shape != coefs[0].shape for c in coefs):
        return None
    return (np.stack(coefs) if coefs else np.zeros((0, 0, 0))), np.array(times, dtype=np.float64)


def segments_to_file(filename, coef, times):
    try:
        with open(filename, "w") as f:
            f.write(segments_to_yaml(coef, times))
    except OSError:
        return False
    return True


def segments_from_file(filename):
    try:
        with open(filename) as f:
            return segments_from_yaml(f.read())
    except OSError:
        return None
