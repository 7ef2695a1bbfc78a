"""Analytic anchors for the two stages whose reference arithmetic is not in the tree (a1 dynamics, a2 LiDAR: PARITY
UNPINNED).  They do not replace a reference trace -- none exists -- but they tie the oracle (and through the
GPU parity tests the kernels) to closed-form results of the model the spec names (single-track / bicycle model, SURVEY.md
Appendix C; wheelbase 0.3302 m [REF ros_agent/agents/follow_the_gap/src/agent.py:78]) and to plain geometry.
"""
import dataclasses

import numpy as np
import pytest

from oracle import Oracle, default_config
from racing_dreamer_b200 import load_track


def _orc(n=1, **kw):
    cfg = default_config()
    cfg.n_envs = n
    for k, v in kw.items():
        setattr(cfg, k, v)
    return Oracle(cfg, [load_track("austria")])


def _state(v=0.0, steer=0.0, yaw=0.0):
    s = np.zeros((7, 1))
    s[2, 0], s[3, 0], s[4, 0] = steer, v, yaw
    return s


def test_straight_line_acceleration_follows_the_first_order_law():
    """steering 0: dv/dt = motor * a_drive - c_drag * v with the drive command sampled once per 10 ms tick (zero-order
    hold, as an ESC would):  v_k = v_inf (1 - (1 - c dt)^k), which tends to v_inf (1 - exp(-c t)) as dt -> 0."""
    orc = _orc()
    p = orc.cfg.vehicle
    motor, T, dt = 0.5, 150, orc.cfg.dt                   # 1.5 s
    s = orc.dynamics(_state(), np.array([[motor, 0.0]]), T)
    v_inf = motor * p.a_drive / p.c_drag
    assert v_inf < p.v_max
    q = 1.0 - p.c_drag * dt
    v = v_inf * (1.0 - q ** T)
    # within a tick the acceleration is constant, so x advances by dt * (v_k + a_k dt / 2)
    vk = v_inf * (1.0 - q ** np.arange(T))
    x = np.sum(dt * (vk + (motor * p.a_drive - p.c_drag * vk) * dt / 2))
    assert s[3, 0] == pytest.approx(v, rel=1e-12)
    assert s[0, 0] == pytest.approx(x, rel=1e-12) and abs(s[1, 0]) < 1e-12 and abs(s[4, 0]) < 1e-12
    assert v == pytest.approx(v_inf * (1.0 - np.exp(-p.c_drag * T * dt)), rel=5e-3)   # the continuous-time law


def test_speed_saturates_at_v_max_and_braking_stops_at_v_min():
    orc = _orc()
    p = orc.cfg.vehicle
    s = orc.dynamics(_state(), np.array([[1.0, 0.0]]), 1500)
    assert s[3, 0] == pytest.approx(p.v_max, abs=0.08) and s[3, 0] <= p.v_max + 0.08   # one tick of overshoot at most
    s = orc.dynamics(s, np.array([[-1.0, 0.0]]), 400)
    assert p.v_min - 0.1 <= s[3, 0] <= p.v_min + 1e-9


def test_low_speed_cornering_is_the_kinematic_bicycle_circle():
    """Kinematic regime (|v| < v_kinematic): yaw rate = v tan(delta) / L, so the car drives a circle of radius
    L / tan(delta) about the point at distance R to its left (delta > 0)."""
    orc = _orc()
    p = orc.cfg.vehicle
    L = p.lf + p.lr
    assert L == pytest.approx(0.3302)                     # [REF ros_agent/agents/follow_the_gap/src/agent.py:78]
    delta, v = 0.3, 0.4
    assert v < p.v_kinematic
    # hold the speed: motor such that motor*a_drive == c_drag*v; steering command that holds delta
    motor = p.c_drag * v / p.a_drive
    steering = delta / (p.steer_gain * p.steer_max)
    T = 300
    s = orc.dynamics(_state(v=v, steer=delta), np.array([[motor, steering]]), T)
    t = T * orc.cfg.dt
    R = L / np.tan(delta)
    yaw = v * t / R
    assert s[3, 0] == pytest.approx(v, rel=1e-9) and s[2, 0] == pytest.approx(delta, abs=1e-12)
    assert s[4, 0] == pytest.approx(yaw, rel=1e-7)
    assert s[0, 0] == pytest.approx(R * np.sin(yaw), rel=1e-6) and s[1, 0] == pytest.approx(R * (1 - np.cos(yaw)), rel=1e-6)


def test_steering_is_rate_limited_and_a_positive_action_turns_right():
    orc = _orc()
    p = orc.cfg.vehicle
    s = orc.dynamics(_state(v=1.0), np.array([[0.2, 1.0]]), 5)          # 50 ms of full right lock
    assert s[2, 0] == pytest.approx(-p.steer_vel_max * 0.05, rel=1e-9)   # steer_gain = -1: negative angle = right
    s = orc.dynamics(s, np.array([[0.2, 1.0]]), 200)
    assert s[2, 0] == pytest.approx(-p.steer_max, abs=1e-9) and s[4, 0] < 0 and s[1, 0] < 0


def test_high_speed_cornering_approaches_the_steady_state_yaw_rate():
    """Dynamic regime, small steering angle: steady state of the linear single-track model,
    yaw_rate = v delta / (L + K v^2) with the understeer gradient from the tyre stiffnesses."""
    orc = _orc()
    p = orc.cfg.vehicle
    L, g = p.lf + p.lr, 9.81
    v, delta = 3.0, 0.02
    motor = p.c_drag * v / p.a_drive
    steering = delta / (p.steer_gain * p.steer_max)
    s = orc.dynamics(_state(v=v, steer=delta), np.array([[motor, steering]]), 400)
    # cornering stiffness per axle [N/rad]: mu * c_s * Fz with the static axle loads
    Cf = p.mu * p.c_sf * p.mass * g * p.lr / L
    Cr = p.mu * p.c_sr * p.mass * g * p.lf / L
    K = p.mass / L * (p.lr / Cf - p.lf / Cr)             # understeer gradient [rad s^2 / m]
    want = v * delta / (L + K * v * v)
    assert s[5, 0] == pytest.approx(want, rel=2e-3)      # small-angle linearisation of the same model
    assert s[3, 0] == pytest.approx(v, rel=1e-3)


def test_lidar_in_a_rectangular_room_matches_the_wall_distances():
    """A 6 m x 3 m room cut into an empty map: every beam must return the analytic distance to the wall it faces, up to
    the grid (the wall is the edge of the first non-drivable cell: within one cell diagonal)."""
    tm = load_track("austria")
    drv = np.zeros_like(tm.drivable)
    r0, c0, hh, ww = 150, 200, 60, 120                   # rows x cols of 0.05 m cells: 3 m x 6 m
    drv[r0:r0 + hh, c0:c0 + ww] = True
    room = dataclasses.replace(tm, drivable=drv)
    cfg = default_config()
    cfg.n_envs = 1
    orc = Oracle(cfg, [room])
    res = tm.resolution
    # world coordinates of the room: crop cell (r, c) -> x = ox + (c0_crop + c) res, y from the flipped row
    x_lo = tm.origin[0] + (tm.c0 + c0) * res
    x_hi = x_lo + ww * res
    y_hi = tm.origin[1] + (tm.full_shape[0] - (tm.r0 + r0)) * res
    y_lo = y_hi - hh * res
    rng = np.random.RandomState(0)
    for _ in range(20):
        x, y = rng.uniform(x_lo + 0.4, x_hi - 0.4), rng.uniform(y_lo + 0.4, y_hi - 0.4)
        yaw = rng.uniform(-np.pi, np.pi)
        got = orc.lidar_cast(np.array([[x, y, yaw]]))[0].astype(np.float64)
        ang = yaw + (0.5 * cfg.lidar_fov - np.arange(1080) * cfg.lidar_fov / 1079.0)   # beam 0 = +135 deg (left)
        dx, dy = np.cos(ang), np.sin(ang)
        with np.errstate(divide="ignore"):
            tx = np.where(dx > 0, (x_hi - x) / dx, np.where(dx < 0, (x_lo - x) / dx, np.inf))
            ty = np.where(dy > 0, (y_hi - y) / dy, np.where(dy < 0, (y_lo - y) / dy, np.inf))
        want = np.clip(np.minimum(tx, ty), cfg.lidar_range_min, cfg.lidar_range_max)
        assert np.abs(got - want).max() < res * 1.5, np.abs(got - want).max()
        assert np.abs(got - want).mean() < res * 0.1      # quantisation of origin (2^-12 cell) and direction (2^-18) only
