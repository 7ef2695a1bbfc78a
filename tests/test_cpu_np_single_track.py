"""The C oracle's dynamics (oracle/rd_oracle.c) against the independent float64 NumPy rendition of the single-track
model (tests/np_single_track.py, textbook form of SURVEY.md Appendix C) -- north_star's dynamics checker.

The two share no code and no operation order (the oracle hoists reciprocals and shares one sine/cosine pair; the NumPy
form divides three times and calls sin / cos / tan separately), so agreement to ~1e-12 after 400 ticks says the
equations are the same; the kernel is then held to both."""
import numpy as np
import pytest

from np_single_track import integrate, params_from_config, random_states, DEFAULT


@pytest.fixture(scope="module")
def orc():
    from oracle import Oracle, default_config
    from racing_dreamer_b200 import load_track
    cfg = default_config()
    cfg.n_envs = 1
    return Oracle(cfg, [load_track("austria")], None, n_threads=1)


def test_defaults_match_the_abi(orc):
    # the NumPy file's own parameter table == rd_default_config's vehicle (include/rd_env.h)
    assert params_from_config(orc.cfg) == DEFAULT


@pytest.mark.parametrize("ticks", [1, 8, 400])
def test_oracle_dynamics_matches_numpy(orc, ticks):
    rng = np.random.RandomState(11)
    state, cmd = random_states(2048, rng)
    p = params_from_config(orc.cfg)
    want = integrate(p, state, cmd, ticks, dt=float(orc.cfg.dt))
    got = orc.dynamics(state, cmd, ticks)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < 1e-9, f"ticks={ticks}: max rel err {err.max():.3e}"


def test_switch_between_kinematic_and_dynamic_regime(orc):
    """States that start just below / above v_kinematic and accelerate or brake through the switch."""
    rng = np.random.RandomState(12)
    state, cmd = random_states(2048, rng, v_lo=0.3, v_hi=0.7)
    state[5] *= 0.2
    state[6] *= 0.2
    p = params_from_config(orc.cfg)
    v0 = state[3].copy()
    want = integrate(p, state, cmd, 60, dt=float(orc.cfg.dt))
    got = orc.dynamics(state, cmd, 60)
    crossed = (v0 < p["v_kinematic"]) != (want[3] < p["v_kinematic"])
    assert crossed.sum() > 200          # the sample really exercises the switch, in both directions
    assert ((v0 < p["v_kinematic"]) & crossed).sum() > 50 and ((v0 >= p["v_kinematic"]) & crossed).sum() > 50
    err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < 1e-9, f"max rel err {err.max():.3e}"


def test_constraints_saturate(orc):
    """Full-lock steering at the stops, full throttle at v_max, braking at standstill: the clipped branches."""
    n = 64
    state = np.zeros((7, n))
    state[2] = np.where(np.arange(n) % 2 == 0, 0.42, -0.42)
    state[3] = np.where(np.arange(n) % 4 < 2, 5.0, 0.0)
    cmd = np.stack([np.where(np.arange(n) % 4 < 2, 1.0, -1.0), np.where(np.arange(n) % 2 == 0, -1.0, 1.0)], 1).astype(np.float64)
    p = params_from_config(orc.cfg)
    want = integrate(p, state, cmd, 50, dt=float(orc.cfg.dt))
    got = orc.dynamics(state, cmd, 50)
    assert np.all(want[3] <= 5.0 + 1e-12) and np.all(want[3] >= -1e-12)
    assert np.all(np.abs(want[2]) <= 0.42 + 1e-12)
    assert np.abs(got - want).max() < 1e-9
