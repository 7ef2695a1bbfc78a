"""Property tests of the CPU oracle (hypothesis) -- the size-independent laws SURVEY.md §8-c(ii) names.  The CUDA path is
held to the same oracle bit for bit / within tolerance in the `-m gpu` tests, so a law that holds here holds there.

* LiDAR: ranges stay in [range_min, range_max]; they never grow when walls are dilated; turning the car by whole beam
  steps shifts the scan; a scan does not depend on what else is in the batch.
* OccupancyMapObs: a heading of exactly 0 makes scipy's rotate the identity, so the observation must be the plain
  220 -> 200 centre crop put through PIL's resize [REF dreamer/wrappers.py:398-406].
* Step: R single-tick steps with the reference's ActionRepeat rule (sum, stop at the first done) == one fused step of
  action_repeat R [REF dreamer/wrappers.py:107-116]; per-tick rewards of maximize_progress telescope to the progress made.
"""
import dataclasses

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import Oracle, default_config
from racing_dreamer_b200 import _abi, load_track

SET = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
TRACKS = ["austria", "columbia", "treitlstrasse_v2"]
_orc = {}


def oracle_for(name, **cfg_kw):
    key = (name, tuple(sorted(cfg_kw.items())))
    if key not in _orc:
        cfg = default_config()
        cfg.n_envs = 4
        for k, v in cfg_kw.items():
            setattr(cfg, k, v)
        _orc[key] = (Oracle(cfg, [load_track(name)], n_threads=2), load_track(name))
    return _orc[key]


def pose_strategy():
    return st.tuples(st.sampled_from(TRACKS), st.integers(0, 10 ** 6), st.floats(-0.2, 0.2), st.floats(-0.2, 0.2),
                     st.floats(-np.pi, np.pi))


def make_pose(tm, idx, dx, dy, yaw):
    p = tm.reset_poses[idx % len(tm.reset_poses)].copy()
    return np.array([p[0] + dx, p[1] + dy, yaw])


@settings(**SET)
@given(pose_strategy())
def test_lidar_bounds_and_batch_independence(args):
    name, idx, dx, dy, yaw = args
    orc, tm = oracle_for(name)
    p = make_pose(tm, idx, dx, dy, yaw)
    alone = orc.lidar_cast(p[None])[0]
    assert alone.min() >= np.float32(orc.cfg.lidar_range_min) and alone.max() <= np.float32(orc.cfg.lidar_range_max)
    others = np.stack([make_pose(tm, idx + 17 * k, dy, dx, yaw + k) for k in range(1, 4)])
    batch = orc.lidar_cast(np.concatenate([others[:2], p[None], others[2:]]))
    assert np.array_equal(batch[2], alone)


@settings(**SET)
@given(pose_strategy(), st.integers(1, 3))
def test_lidar_never_grows_when_walls_are_dilated(args, r):
    name, idx, dx, dy, yaw = args
    orc, tm = oracle_for(name)
    from scipy import ndimage
    thin = ndimage.binary_erosion(tm.drivable, structure=np.ones((3, 3), bool), iterations=r)   # walls r cells thicker
    p = make_pose(tm, idx, dx, dy, yaw)
    row, col = tm.to_pixel(p[0], p[1])
    if not thin[row - tm.r0, col - tm.c0]:
        return                                   # the sensor itself would sit inside the thicker wall
    tm2 = dataclasses.replace(tm, drivable=thin)
    cfg = default_config()
    cfg.n_envs = 1
    a = orc.lidar_cast(p[None])[0]
    b = Oracle(cfg, [tm2]).lidar_cast(p[None])[0]
    assert np.all(b <= a)
    # and by no more than the dilation can explain along any beam that still sees a wall in range
    # (a wall r cells thicker is hit at most ~r cells / sin(incidence) earlier: only the sign is a law)


@settings(**SET)
@given(pose_strategy(), st.integers(1, 40))
def test_turning_by_whole_beams_shifts_the_scan(args, k):
    name, idx, dx, dy, yaw = args
    orc, tm = oracle_for(name)
    p = make_pose(tm, idx, dx, dy, yaw)
    step = float(orc.cfg.lidar_fov) / (orc.nb - 1)
    q = p.copy()
    q[2] -= k * step                             # turned right by k beams: beam i now looks where beam i + k did
    a, b = orc.lidar_cast(p[None])[0], orc.lidar_cast(q[None])[0]
    d = np.abs(b[:-k] - a[k:])
    # the ray is re-quantised (direction to 2^-18), so grazing beams may flip to the next wall: a law of the median
    assert np.median(d) < 1e-3 and (d < 0.1).mean() > 0.9


@settings(**SET)
@given(st.sampled_from(TRACKS), st.integers(0, 10 ** 6))
def test_occupancy_at_heading_zero_is_crop_and_resize(name, idx):
    Image = pytest.importorskip("PIL.Image")
    orc, tm = oracle_for(name)
    p = tm.reset_poses[idx % len(tm.reset_poses)].copy()
    p[2] = 0.0
    got = orc.occupancy_obs(p[None])[0]
    full = tm.full_drivable()
    pr, pc = tm.to_pixel(p[0], p[1])
    crop = full[pr - 110:pr + 110, pc - 110:pc + 110].astype(np.uint8)     # [REF dreamer/wrappers.py:398-400]
    cr, cc = crop.shape[0] // 2, crop.shape[1] // 2
    mid = crop[cr - 100:cr + 100, cc - 100:cc + 100]                      # rotate(., 360 deg) == identity
    want = np.array(Image.fromarray(mid).resize(size=(64, 64)))
    assert np.array_equal(got, want)


@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.sampled_from(TRACKS), st.integers(0, 2 ** 31 - 1), st.sampled_from([2, 4, 8]))
def test_fused_action_repeat_equals_single_ticks(name, seed, R):
    fused, tm = oracle_for(name, action_repeat=R, auto_reset=0, n_envs=4)
    tick, _ = oracle_for(name, action_repeat=1, auto_reset=0, n_envs=4)
    rng = np.random.RandomState(seed)
    for o in (fused, tick):
        o.cfg.seed = seed & 0xFFFF
        o.i32[:] = 0                             # the oracles are cached across examples: same episode counters
        o.f64[:] = 0.0
        o.reset(mode=_abi.RESET_RANDOM)
    for step in range(12):
        a = (rng.uniform(-1, 1, (4, 2)) * [1.0, 0.4]).astype(np.float32)
        before = fused.i32[_abi.I_LAP] + fused.f64[_abi.S_PROGRESS]
        f = {k: v.copy() for k, v in fused.step(a).items() if v is not None}
        total, done = np.zeros(4), np.zeros(4, bool)
        for _ in range(R):                       # ActionRepeat [REF dreamer/wrappers.py:107-116], per env
            live = ~done
            tick.i32[_abi.I_FLAGS, done] |= _abi.F_NEEDS_RESET          # a done env takes no more ticks
            tick.i32[_abi.I_FLAGS, live] &= ~_abi.F_NEEDS_RESET
            t = tick.step(a)
            total[live] += t["reward64"][live]
            done[live] = t["done"][live].astype(bool)
        assert np.array_equal(done, f["done"].astype(bool))
        assert np.allclose(total, f["reward64"], rtol=0, atol=1e-12)
        assert np.array_equal(tick.f64[:7], fused.f64[:7])
        # maximize_progress: the step's reward is 100 x the progress made (+ the collision penalty)
        after = fused.i32[_abi.I_LAP] + fused.f64[_abi.S_PROGRESS]
        pen = np.where((f["flags"] & _abi.F_COLLISION) != 0, -1.0, 0.0)
        wrapped = np.abs(after - before) > 0.5
        assert np.allclose(f["reward64"][~wrapped], (100.0 * (after - before) + pen)[~wrapped], atol=1e-9)
        if done.any():
            for o in (fused, tick):
                o.reset(mask=done.astype(np.uint8), mode=_abi.RESET_RANDOM)
