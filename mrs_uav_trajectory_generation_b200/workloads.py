"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d).

`random_flier_paths` follows the distribution of the reference's test path generator (src/path_random_flier.cpp:317-351
with the constants of tmux/dynamic_test/config/path_random_flier.yaml:5-33): start (0, 0, 5), bearing0 ~ U(-pi, pi),
per step bearing += U(-0.4, 0.4), distance ~ U(0.5, 2.0) m, z = 5 + U(-0.1, 0.1), waypoint heading = bearing.
A counter-based generator (numpy Philox keyed by 0xB200 + problem index) replaces the reference's rand().
"""
import numpy as np

SEED_BASE = 0xB200


def random_flier_path(index, n_waypoints=11):
    rng = np.random.Generator(np.random.Philox(key=SEED_BASE + int(index)))
    bearing = rng.uniform(-np.pi, np.pi)
    x = y = 0.0
    out = np.empty((n_waypoints, 4))
    for i in range(n_waypoints):
        bearing += rng.uniform(-0.4, 0.4)
        dist = rng.uniform(0.5, 2.0)
        x += np.cos(bearing) * dist
        y += np.sin(bearing) * dist
        z = 5.0 + rng.uniform(-0.1, 0.1)
        out[i] = (x, y, z, bearing)
    return out


def random_flier_paths(B, n_waypoints=11, first_index=0):
    """Returns (wp_off int32[B+1], wp float64[B*n_waypoints, 4]).  Vectorised over problems, same stream per problem
    as random_flier_path."""
    wp = np.empty((B, n_waypoints, 4))
    for p in range(B):
        wp[p] = random_flier_path(first_index + p, n_waypoints)
    wp_off = (np.arange(B + 1) * n_waypoints).astype(np.int32)
    return wp_off, wp.reshape(-1, 4)


def random_flier_paths_fast(B, n_waypoints=11, first_index=0):
    """Bulk variant for large batches (one Philox stream keyed by first_index, jumped per problem is too slow in Python):
    draws all uniforms from a single keyed stream.  Used by bench.py for 65k+ problems; parity tests use the per-problem
    generator above."""
    rng = np.random.Generator(np.random.Philox(key=SEED_BASE + (int(first_index) << 20) + 1))
    b0 = rng.uniform(-np.pi, np.pi, size=(B, 1))
    db = rng.uniform(-0.4, 0.4, size=(B, n_waypoints))
    dist = rng.uniform(0.5, 2.0, size=(B, n_waypoints))
    dz = rng.uniform(-0.1, 0.1, size=(B, n_waypoints))
    bearing = b0 + np.cumsum(db, axis=1)
    x = np.cumsum(np.cos(bearing) * dist, axis=1)
    y = np.cumsum(np.sin(bearing) * dist, axis=1)
    wp = np.stack([x, y, 5.0 + dz, bearing], axis=2)
    wp_off = (np.arange(B + 1) * n_waypoints).astype(np.int32)
    return wp_off, np.ascontiguousarray(wp.reshape(-1, 4))


# SURVEY.md 8(d) config 1 fixtures
F1A_WAYPOINTS = np.array([[10, 20, 3.5, 1.2], [-5, -5, 5, 1], [-5, 5, 5, 2], [5, -5, 5, 3], [5, 5, 5, 4]], dtype=np.float64)
F1A_INIT_HEADING = 1.2
F1B_WAYPOINTS = np.array([[0, 0, 3, 0]] + [[2.0 * i, 0.5 if i % 2 == 0 else -0.5, 5, 0] for i in range(10)], dtype=np.float64)
F1B_INIT_HEADING = 0.0


def init14(heading, vel=(0, 0, 0, 0), acc=(0, 0, 0, 0), jerk=(0, 0, 0, 0)):
    return np.array([1.0, heading, *vel, *acc, *jerk], dtype=np.float64)
