"""Independent float64 NumPy rendition of the single-track vehicle model (north_star:
"compared against a float64 NumPy rendition of the same model").

TEST INFRASTRUCTURE.  Written from the textbook (CommonRoad / f1tenth single-track) equations of Appendix C, NOT from
oracle/rd_oracle.c or the CUDA kernel: three true divisions, separate numpy sin / cos / tan calls, no shared
sub-expressions, the classical RK4 formula.  Rule: this file never follows an optimisation of the
kernel or of the C oracle -- if they drift away from it, they are wrong.

State q = (x, y, delta, v, psi, psi_dot, beta), all arrays of shape [n].
Inputs per 10 ms tick: `motor` and `steering`, the sim-facing command [REF dreamer/dream.py:138 motor in [0.005, 1],
steering in [-1, 1] after ReduceActionSpace].
"""
import numpy as np

G = 9.81

# the f1tenth parameter set of Appendix C with the in-tree anchors (wheelbase 0.3302 = lf + lr
# [REF ros_agent/agents/follow_the_gap/src/agent.py:78], max steering 0.42 and max speed 5
# [REF ros_agent/models/dreamer/racing_dreamer.py:14-16], max deceleration 8.26 [REF agent.py:74])
DEFAULT = dict(mu=1.0489, c_sf=4.718, c_sr=5.4562, lf=0.15875, lr=0.17145, h_cg=0.074, mass=3.74, inertia=0.04712,
               steer_min=-0.42, steer_max=0.42, steer_vel_max=3.2, v_switch=7.319, a_max=9.51, v_min=0.0, v_max=5.0,
               v_kinematic=0.5, a_drive=6.0, a_brake=8.26, c_drag=1.0, steer_gain=-1.0)


def params_from_config(cfg):
    """dict of the vehicle parameters held by an rd_config (ctypes mirror of include/rd_env.h)."""
    v = cfg.vehicle
    return {k: float(getattr(v, k)) for k in DEFAULT}


def steering_rate(p, delta, rate):
    """CommonRoad steering constraint: no motion into a stop, rate clipped to +-steer_vel_max."""
    out = np.clip(rate, -p["steer_vel_max"], p["steer_vel_max"])
    blocked = ((delta <= p["steer_min"]) & (rate <= 0.0)) | ((delta >= p["steer_max"]) & (rate >= 0.0))
    return np.where(blocked, 0.0, out)


def acceleration(p, v, acc):
    """CommonRoad acceleration constraint: power-limited above v_switch, no acceleration beyond the speed limits."""
    with np.errstate(divide="ignore", invalid="ignore"):
        upper = np.where(v > p["v_switch"], p["a_max"] * p["v_switch"] / v, p["a_max"])
    out = np.minimum(np.maximum(acc, -p["a_max"]), upper)
    blocked = ((v <= p["v_min"]) & (acc <= 0.0)) | ((v >= p["v_max"]) & (acc >= 0.0))
    return np.where(blocked, 0.0, out)


def rhs(p, q, rate_cmd, acc_cmd):
    """dq/dt of the single-track model; kinematic bicycle below v_kinematic."""
    x, y, delta, v, psi, psi_dot, beta = q
    mu, c_sf, c_sr = p["mu"], p["c_sf"], p["c_sr"]
    lf, lr, h, m, iz = p["lf"], p["lr"], p["h_cg"], p["mass"], p["inertia"]
    lwb = lf + lr
    d_delta = steering_rate(p, delta, rate_cmd)
    a = acceleration(p, v, acc_cmd)
    with np.errstate(divide="ignore", invalid="ignore"):
        # ---- dynamic single-track (Appendix C) ----
        dyn = [
            v * np.cos(psi + beta),
            v * np.sin(psi + beta),
            d_delta,
            a,
            psi_dot,
            -mu * m / (v * iz * lwb) * (lf ** 2 * c_sf * (G * lr - a * h) + lr ** 2 * c_sr * (G * lf + a * h)) * psi_dot
            + mu * m / (iz * lwb) * (lr * c_sr * (G * lf + a * h) - lf * c_sf * (G * lr - a * h)) * beta
            + mu * m / (iz * lwb) * lf * c_sf * (G * lr - a * h) * delta,
            (mu / (v ** 2 * lwb) * (c_sr * (G * lf + a * h) * lr - c_sf * (G * lr - a * h) * lf) - 1.0) * psi_dot
            - mu / (v * lwb) * (c_sr * (G * lf + a * h) + c_sf * (G * lr - a * h)) * beta
            + mu / (v * lwb) * c_sf * (G * lr - a * h) * delta,
        ]
    # ---- kinematic bicycle: psi_dot = v / lwb * tan(delta); its time derivative by the product rule ----
    kin = [
        v * np.cos(psi),
        v * np.sin(psi),
        d_delta,
        a,
        v / lwb * np.tan(delta),
        a / lwb * np.tan(delta) + v / (lwb * np.cos(delta) ** 2) * d_delta,
        np.zeros_like(v),
    ]
    slow = np.abs(v) < p["v_kinematic"]
    return [np.where(slow, k, d) for k, d in zip(kin, dyn)]


def tick(p, q, motor, steering, dt=0.01):
    """One RK4 step of dt under the command (motor, steering), held constant over the step."""
    delta, v = q[2], q[3]
    target = steering * p["steer_gain"] * p["steer_max"]
    rate_cmd = (target - delta) / dt
    drive = np.where(motor >= 0.0, motor * p["a_drive"], motor * p["a_brake"])
    acc_cmd = drive - p["c_drag"] * v
    k1 = rhs(p, q, rate_cmd, acc_cmd)
    k2 = rhs(p, [a + 0.5 * dt * b for a, b in zip(q, k1)], rate_cmd, acc_cmd)
    k3 = rhs(p, [a + 0.5 * dt * b for a, b in zip(q, k2)], rate_cmd, acc_cmd)
    k4 = rhs(p, [a + dt * b for a, b in zip(q, k3)], rate_cmd, acc_cmd)
    return [a + dt / 6.0 * (b + 2.0 * c + 2.0 * d + e) for a, b, c, d, e in zip(q, k1, k2, k3, k4)]


def integrate(p, state, commands, n_ticks, dt=0.01):
    """state [7, n] float64, commands [n, 2] = (motor, steering) -> state after n_ticks ticks."""
    q = [np.array(state[i], dtype=np.float64) for i in range(7)]
    motor = np.asarray(commands[:, 0], np.float64)
    steering = np.asarray(commands[:, 1], np.float64)
    for _ in range(int(n_ticks)):
        q = tick(p, q, motor, steering, dt)
    return np.stack(q)


def random_states(n, rng, v_lo=0.0, v_hi=4.5):
    """The state distribution of the parity tests: speeds straddle the kinematic/dynamic switch at 0.5 m/s."""
    s = np.zeros((7, n))
    s[0] = rng.uniform(-5, 5, n)
    s[1] = rng.uniform(-5, 5, n)
    s[2] = rng.uniform(-0.4, 0.4, n)
    s[3] = rng.uniform(v_lo, v_hi, n)
    s[4] = rng.uniform(-np.pi, np.pi, n)
    s[5] = rng.uniform(-1, 1, n)
    s[6] = rng.uniform(-0.2, 0.2, n)
    cmd = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)], 1)
    return s, cmd
