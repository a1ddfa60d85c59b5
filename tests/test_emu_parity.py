"""CPU-side parity: the product's device code + host pipeline, compiled for the host by tests/host_emu (same C ABI),
against the oracle.  These run without a GPU; tests/test_gpu_parity.py repeats them through libtg_b200.so on the B200."""
import numpy as np
import pytest

import parity_checks as PC


@pytest.mark.parametrize("r", [2, 3, 4])
def test_linear_batch(emu_ctx, oracle, r):
    worst, exact = PC.check_linear_batch(emu_ctx, seed=10 + r, B=24, r=r)
    assert exact, worst


@pytest.mark.parametrize("r", [2, 3])
def test_time_alloc_from_vertices(emu_ctx, oracle, r):
    assert PC.check_time_alloc(emu_ctx, r=r)


def test_preprocess_fallback_and_waypoint_indices(emu_ctx, oracle):
    assert PC.check_path_side_steps(emu_ctx)


def test_scaling_certificates_and_multi_pass(emu_ctx, oracle):
    assert PC.check_scaling_multi_pass(emu_ctx)


def test_sampling(emu_ctx, oracle):
    assert PC.check_sampling(emu_ctx)


def test_evaluate(emu_ctx, oracle):
    assert PC.check_evaluate(emu_ctx)


def test_objectives_of_the_other_time_allocation_methods(emu_ctx, oracle):
    assert PC.check_objectives(emu_ctx)


def test_derivative_free_time_allocation(emu_ctx, oracle):
    assert PC.check_derivative_free_time_allocation(emu_ctx)


def test_two_lanes(emu_lib, oracle):
    assert PC.check_two_lanes(emu_lib, n=48)


def test_max_magnitude(emu_ctx, oracle):
    assert PC.check_max_magnitude(emu_ctx)


def test_extrema_and_scaling(emu_ctx, oracle):
    assert PC.check_extrema_and_scaling(emu_ctx)


def test_random_flier_full_pipeline(emu_ctx, oracle):
    res, out, exact, worst = PC.check_random_flier(emu_ctx, 48)
    assert exact
    assert res["success"].all()


def test_fixtures(emu_ctx, oracle):
    res, out, exact = PC.check_fixtures(emu_ctx)
    assert exact
    # the reference tests' geometric acceptance predicate on the original waypoints (get_path_test.h:45-68)
    from mrs_uav_trajectory_generation_b200 import workloads as W

    for p, wps in enumerate([W.F1A_WAYPOINTS, W.F1B_WAYPOINTS]):
        smp = out["samples"][out["smp_off"][p]:out["smp_off"][p + 1]]
        assert PC.geometric_predicate(smp, wps[1:])


def test_mixed_ragged_batch(emu_ctx, oracle):
    res, out, exact, worst = PC.check_mixed_batch(emu_ctx)
    assert exact


@pytest.mark.parametrize("r", [2, 4])
def test_config2_linear_plus_sampling(emu_ctx, oracle, r):
    res, out, exact, worst = PC.check_config2(emu_ctx, B=32, r=r)
    assert exact


def test_sweep(emu_ctx, oracle):
    assert PC.check_sweep(emu_ctx, K=200)


def test_jerk_and_snap_full_pipeline(emu_ctx, oracle):
    for r in (3, 4):
        res, out, exact, worst = PC.check_random_flier(emu_ctx, 6, first_index=300, derivative_to_optimize=r)
        assert exact


def test_extrema_random_polynomials(emu_ctx, oracle):
    import oracle_lib as O

    rng = np.random.default_rng(4)
    S = 1500
    coef = rng.standard_normal((S, 4, 10)) * np.exp(rng.uniform(-6, 3, (S, 1, 10)))
    coef[::7, :, 1:3] = 0.0
    coef[::11, :, 9] = 0.0
    coef[::13, :, 8:] = 0.0
    times = np.exp(rng.uniform(-2, 2, S))
    assert np.array_equal(emu_ctx.extrema(coef, times), O.segment_maxima(coef, times))


def test_acceptance_reject_branches(emu_ctx, oracle):
    assert PC.check_acceptance_rejects(emu_ctx, B=24)


def test_device_jenkins_traub_against_reference_vectors(emu_ctx, oracle):
    assert PC.check_roots_against_reference_vectors(emu_ctx) == 600
    assert PC.check_roots_adversarial(emu_ctx, n=300)


def test_override_heading_atan2(emu_ctx, oracle):
    assert PC.check_heading_override(emu_ctx)


def test_sweep_best_over_several_contexts(emu_ctx, oracle):
    assert PC.check_sweep_best(emu_ctx)


def test_degenerate_inputs(emu_ctx, oracle):
    assert PC.check_degenerate_inputs(emu_ctx)


def test_batch_cut_into_several_groups(emu_lib, oracle):
    assert PC.check_several_groups(emu_lib)
