"""The mrs_msgs::Path override fields (node.cpp:1847-1902) and what findTrajectory does with them (node.cpp:972-1040): host logic
checked against hand-written expectations on CPU; the batch call they drive is checked against the oracle on the emulator / GPU."""
import numpy as np
import pytest

import oracle_lib as O
from mrs_uav_trajectory_generation_b200 import api, workloads as W

CONSTRAINTS = dict(horizontal_speed=4.0, vertical_ascending_speed=2.5, vertical_descending_speed=2.0, horizontal_acceleration=3.0,
                   vertical_ascending_acceleration=2.0, vertical_descending_acceleration=1.5, horizontal_jerk=30.0, vertical_ascending_jerk=25.0,
                   vertical_descending_jerk=20.0, heading_speed=1.0, heading_acceleration=2.0, heading_jerk=10.0)


class _P:  # the two fields resolve_request reads from the base parameters
    max_deviation = 0.2


def _points(n=5, seed=0):
    return W.random_flier_path(seed, n)


def test_plain_request_takes_the_smaller_vertical_limits():
    c = api.DynamicsConstraints(**CONSTRAINTS)
    wp, stop, L, max_dev, prepend, over = api.resolve_request(api.PathRequest(_points()), c, _P)
    assert L == [4.0, 2.0, 3.0, 1.5, 30.0, 20.0, 1.0, 2.0, 10.0]
    assert max_dev == 0.2 and not prepend and not over and not stop.any() and len(wp) == 5


def test_loop_stop_relax_and_deviation_fields():
    c = api.DynamicsConstraints(**CONSTRAINTS)
    req = api.PathRequest(_points(), loop=True, stop_at_waypoints=True, relax_heading=True, max_deviation_from_path=0.5)
    wp, stop, L, max_dev, prepend, over = api.resolve_request(req, c, _P)
    assert len(wp) == 6 and np.array_equal(wp[-1], wp[0]) and stop.all()
    assert L[6:] == [api.FLT_MAX] * 3 and L[:6] == [4.0, 2.0, 3.0, 1.5, 30.0, 20.0]
    assert max_dev == 0.5


def test_override_uses_the_horizontal_jerk_for_the_vertical_axis():
    c = api.DynamicsConstraints(**CONSTRAINTS)
    req = api.PathRequest(_points(), override_constraints=True, override_max_velocity_horizontal=8.0, override_max_acceleration_horizontal=4.0,
                          override_max_jerk_horizontal=40.0, override_max_velocity_vertical=3.0, override_max_acceleration_vertical=2.5,
                          override_max_jerk_vertical=99.0)
    *_, L, max_dev, prepend, over = api.resolve_request(req, c, _P)
    assert over and L[:6] == [8.0, 3.0, 4.0, 2.5, 40.0, 40.0]  # node.cpp:1856: jerk_vertical_ := override_max_jerk_horizontal


def test_override_refused_when_the_current_state_is_beyond_it():
    c = api.DynamicsConstraints(**CONSTRAINTS)
    kw = dict(override_constraints=True, override_max_velocity_horizontal=1.0, override_max_acceleration_horizontal=4.0, override_max_jerk_horizontal=40.0,
              override_max_velocity_vertical=3.0, override_max_acceleration_vertical=2.5)
    fast = W.init14(heading=0.3, vel=(0.8, 0.8, 0.0, 0.0))  # |v_xy| = 1.13 > 1.0
    *_, L, _, prepend, over = api.resolve_request(api.PathRequest(_points(), **kw), c, _P, fast)
    assert prepend and not over and L[:6] == [4.0, 2.0, 3.0, 1.5, 30.0, 20.0]
    slow = W.init14(heading=0.3, vel=(0.5, 0.5, 0.0, 0.0))
    *_, L, _, prepend, over = api.resolve_request(api.PathRequest(_points(), **kw), c, _P, slow)
    assert over and L[0] == 1.0
    # dont_prepend_current_state: no initial state reaches findTrajectory, so nothing can refuse the override (node.cpp:508-510, 1002)
    *_, L, _, prepend, over = api.resolve_request(api.PathRequest(_points(), dont_prepend_current_state=True, **kw), c, _P, fast)
    assert not prepend and over and L[0] == 1.0


def _check_requests(ctx):
    gen = api.TrajectoryGenerator(ctx)
    c = api.DynamicsConstraints(**CONSTRAINTS)
    over = dict(override_constraints=True, override_max_velocity_horizontal=6.0, override_max_acceleration_horizontal=3.5, override_max_jerk_horizontal=35.0,
                override_max_velocity_vertical=3.0, override_max_acceleration_vertical=2.5)
    reqs, states = [], []
    for i in range(12):
        kw = {}
        if i % 3 == 1:
            kw.update(over)
        if i % 4 == 2:
            kw.update(relax_heading=True, max_deviation_from_path=0.35)
        if i % 5 == 3:
            kw.update(loop=True, stop_at_waypoints=True)
        if i % 6 == 5:
            kw.update(dont_prepend_current_state=True)
        reqs.append(api.PathRequest(_points(5 + i % 4, seed=900 + i), **kw))
        states.append(W.init14(heading=reqs[-1].points[0, 3], vel=(0.4, -0.3, 0.1, 0.0)))
    placed, resolved = gen.optimize_requests(reqs, c, states)
    assert len({id(br) for br, _ in placed}) > 2  # several parameter groups
    for i, (br, k) in enumerate(placed):
        wp, stop, L, max_dev, prepend, _ = resolved[i]
        P = O.default_params(limits=L, max_deviation=max_dev)
        ref = O.optimize_batch(np.array([0, len(wp)], np.int32), wp, stop_at=stop, init=np.asarray([states[i]]) if prepend else None, params=P,
                               cap_wp=400, cap_samples=4000)
        r = ref["res"][0]
        g = br.results[k]
        for f in ("status", "success", "nlopt_code", "n_evals", "rounds", "safe", "n_waypoints", "n_samples"):
            assert getattr(r, f) == g[f], (i, f)
        S, M = r.n_waypoints - 1, r.n_samples
        assert np.array_equal(br.trajectory(k).times, ref["times"][0, :S])
        assert np.array_equal(br.samples(k), ref["samples"][0, :M])
    return True


def test_requests_through_the_batch_call_on_host_emulation(emu_ctx, oracle):
    assert _check_requests(emu_ctx)


@pytest.mark.gpu
def test_requests_through_the_batch_call_on_gpu(gpu_ctx, oracle):
    assert _check_requests(gpu_ctx)
