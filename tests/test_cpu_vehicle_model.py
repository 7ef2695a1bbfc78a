"""Host-compiled check of the kernel's vehicle model (racing_dreamer_b200/csrc/rd_vehicle.cuh, the functions k_step
calls) against BOTH checkers: the C oracle and the independent NumPy rendition of SURVEY.md Appendix C.  The header is
compiled with g++ (tests/native/vehicle_check.cpp): a unit test of the model's arithmetic, not a CPU path of the
product -- the GPU tests then hold the compiled kernel to the same bars."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from np_single_track import integrate, params_from_config, random_states

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("vehicle") / "libvehicle_check.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(out),
                    str(ROOT / "tests" / "native" / "vehicle_check.cpp")], check=True)
    lib = C.CDLL(str(out))
    lib.vehicle_ticks.restype = None
    lib.vehicle_ticks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.vehicle_sincos.restype = None
    lib.vehicle_sincos.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.vehicle_div_mismatches.restype = C.c_longlong
    lib.vehicle_div_mismatches.argtypes = [C.c_int]
    return lib


@pytest.fixture(scope="module")
def orc():
    from oracle import Oracle, default_config
    from racing_dreamer_b200 import load_track
    cfg = default_config()
    cfg.n_envs = 1
    return Oracle(cfg, [load_track("austria")], None, n_threads=1)


def _run(lib, cfg, state, cmd, ticks):
    assert lib.vehicle_sizeof_config() == C.sizeof(cfg)
    s = np.ascontiguousarray(state, dtype=np.float64).copy()
    c = np.ascontiguousarray(cmd, dtype=np.float64)
    lib.vehicle_ticks(C.addressof(cfg), s.ctypes.data, c.ctypes.data, s.shape[1], ticks)
    return s


@pytest.mark.parametrize("ticks", [1, 8, 400])
def test_kernel_model_matches_oracle_and_numpy(lib, orc, ticks):
    state, cmd = random_states(2048, np.random.RandomState(4))
    got = _run(lib, orc.cfg, state, cmd, ticks)
    for name, want in (("oracle", orc.dynamics(state, cmd, ticks)),
                       ("numpy", integrate(params_from_config(orc.cfg), state, cmd, ticks, dt=float(orc.cfg.dt)))):
        err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
        assert err.max() < 1e-9, f"{name}, ticks={ticks}: max rel err {err.max():.3e}"


def test_kernel_model_regime_switch_and_stops(lib, orc):
    rng = np.random.RandomState(12)
    state, cmd = random_states(2048, rng, v_lo=0.3, v_hi=0.7)
    # saturated cases: steering at the stops, speed at the limits, braking at standstill
    state[2, :64] = np.where(np.arange(64) % 2 == 0, 0.42, -0.42)
    state[3, :64] = np.where(np.arange(64) % 4 < 2, 5.0, 0.0)
    cmd[:64, 0] = np.where(np.arange(64) % 4 < 2, 1.0, -1.0)
    cmd[:64, 1] = np.where(np.arange(64) % 2 == 0, -1.0, 1.0)
    got = _run(lib, orc.cfg, state, cmd, 60)
    want = integrate(params_from_config(orc.cfg), state, cmd, 60, dt=float(orc.cfg.dt))
    err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < 1e-9, f"max rel err {err.max():.3e}"
    assert np.abs(got - orc.dynamics(state, cmd, 60)).max() < 1e-9


def test_kernel_model_fast_vehicle_power_limit(lib):
    """A vehicle that exceeds v_switch exercises the power-limited acceleration branch (has_switch)."""
    from oracle import Oracle, default_config
    from racing_dreamer_b200 import load_track
    cfg = default_config()
    cfg.n_envs = 1
    cfg.vehicle.v_max = 12.0
    cfg.vehicle.a_drive = 14.0
    o = Oracle(cfg, [load_track("austria")], None, n_threads=1)
    state, cmd = random_states(512, np.random.RandomState(5), v_lo=5.0, v_hi=11.0)
    cmd[:, 0] = np.abs(cmd[:, 0])
    got = _run(lib, o.cfg, state, cmd, 100)
    want = integrate(params_from_config(o.cfg), state, cmd, 100, dt=float(o.cfg.dt))
    assert (want[3] > 7.319).sum() > 100
    err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < 1e-9, f"max rel err {err.max():.3e}"
    assert np.abs(got - o.dynamics(state, cmd, 100)).max() < 1e-9


def test_sincos_accuracy(lib):
    """<= 2 ulp up to the 1e5 rad guard (fdlibm kernels on a one-word reduced argument: the reduction's tail is dropped)."""
    rng = np.random.RandomState(1)
    x = np.concatenate([rng.uniform(-4, 4, 20000), rng.uniform(-1e5, 1e5, 20000), rng.uniform(-1e-3, 1e-3, 1000),
                        np.array([0.0, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi, 99999.9, -99999.9, 1e6, -3e7])])
    out = np.empty((len(x), 2))
    lib.vehicle_sincos(np.ascontiguousarray(x).ctypes.data, len(x), out.ctypes.data)
    ls, lc = np.longdouble(x), np.longdouble(x)
    es = np.abs(out[:, 0] - np.sin(ls)) / np.spacing(np.abs(np.sin(x)))
    ec = np.abs(out[:, 1] - np.cos(lc)) / np.spacing(np.abs(np.cos(x)))
    assert float(es.max()) <= 2.0 and float(ec.max()) <= 2.0, (float(es.max()), float(ec.max()))


def test_progress_division_is_correctly_rounded(lib):
    """progress = dist / dmax through the reciprocal (Markstein) equals the true division for every value a map can
    hold -- the quotient decides the integer checkpoint index."""
    from racing_dreamer_b200 import available_tracks, load_track
    dmaxes = sorted({int(load_track(t).dmax) for t in available_tracks()})
    assert dmaxes
    for dmax in dmaxes + [1, 2, 3, 7, 1000, 65535]:
        assert lib.vehicle_div_mismatches(dmax) == 0, dmax
