"""CPU checks of the multi-agent worlds in the oracle: the dict-of-agents semantics of the reference's
wrappers [REF dreamer/wrappers.py:107-116,147-154; dreamer/tools.py:178-179] restated per world, car-car contact, scans
that see the other cars, rank, the n_step_progress task and the random_ball reset.  The CUDA path is held to the same
oracle in tests/test_gpu_multi_agent.py.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import Oracle
from oracle.binding import default_config
from racing_dreamer_b200 import _abi
from racing_dreamer_b200.maps import load_track


def make(track="austria", worlds=8, A=4, tasks=None, **kw):
    cfg = default_config()
    cfg.n_envs = worlds * A
    cfg.agents_per_world = A
    for a in range(_abi.MAX_AGENTS):
        cfg.agent_task[a] = (tasks[a] if tasks and a < len(tasks) else _abi.TASK_MAX_PROGRESS)
    cfg.action_repeat = 4
    cfg.auto_reset = 1
    cfg.reset_mode = _abi.RESET_RANDOM_BALL
    cfg.seed = 5
    for k, v in kw.items():
        setattr(cfg, k, v)
    tm = load_track(track)
    return Oracle(cfg, [tm], np.zeros(cfg.n_envs, np.int32), n_threads=2), tm


def test_config_layout_matches_header():
    lib = Oracle.__init__.__globals__["_load"]()
    assert lib.orc_sizeof_config() == C.sizeof(_abi.RdConfig)


def test_grid_reset_uses_the_staggered_slots():
    orc, tm = make(worlds=2, A=4)
    out = orc.reset(mode=_abi.RESET_GRID)
    for w in range(2):
        for a in range(4):
            e = w * 4 + a
            assert np.allclose(orc.f64[[_abi.S_X, _abi.S_Y, _abi.S_YAW], e], tm.start_poses[a])
    assert np.array_equal(out["rank"], np.tile(np.arange(1, 5), 2))


def test_random_ball_keeps_the_cars_of_a_world_close_and_apart():
    orc, tm = make(worlds=64, A=4)
    orc.reset(mode=_abi.RESET_RANDOM_BALL)
    xy = orc.f64[[_abi.S_X, _abi.S_Y]].T.reshape(64, 4, 2)
    d01 = np.linalg.norm(xy[:, 0] - xy[:, 1], axis=1)
    # consecutive cars: about ball_spacing metres of track apart, never overlapping (body diagonal 0.57 m)
    assert np.all(d01 > 0.6) and np.median(d01) < 3.0
    # worlds draw different anchors
    assert len({tuple(np.round(p, 3)) for p in xy[:, 0]}) > 32
    # no car starts in contact with another or with a wall
    out = orc.step(np.zeros((256, 2), np.float32))
    assert not out["opponents"].any() and not (out["flags"] & _abi.F_COLLISION).any()


def test_scans_see_the_other_cars():
    orc, tm = make(worlds=1, A=2, lidar_offset=0.0)
    orc4, _ = make(worlds=1, A=1)
    # car 1 two metres straight ahead of car 0 on the start straight
    x, y, yaw = tm.start_poses[0]
    p0 = np.array([x, y, yaw])
    p1 = np.array([x + 2.0 * np.cos(yaw), y + 2.0 * np.sin(yaw), yaw])
    two = orc.lidar_cast(np.stack([p0, p1]))
    alone = orc4.lidar_cast(np.stack([p0, p1]))
    mid = 1080 // 2
    hl = 0.5 * orc.cfg.vehicle.body_length
    assert abs(two[0, mid] - (2.0 - hl)) < 0.02 and alone[0, mid] > 2.5       # sees the rear bumper
    assert np.array_equal(two[1], alone[1]) or np.all(two[1] <= alone[1])     # car 1 looks away: 270 deg fov still may catch it
    changed = np.nonzero(two[0] != alone[0])[0]
    assert changed.min() > mid - 60 and changed.max() < mid + 60              # only the beams through the car change
    assert np.all(two[0] <= alone[0])


def test_contact_stops_the_world_and_resets_it_together():
    orc, tm = make(worlds=2, A=2, auto_reset=1, reset_mode=_abi.RESET_GRID)
    orc.reset(mode=_abi.RESET_GRID)
    # world 0: put car 1 right in front of car 0 (0.45 m < body length 0.5): contact on the first tick
    x, y, yaw = tm.start_poses[0]
    orc.f64[_abi.S_X, 1] = x + 0.45 * np.cos(yaw)
    orc.f64[_abi.S_Y, 1] = y + 0.45 * np.sin(yaw)
    orc.f64[_abi.S_YAW, 1] = yaw
    ep0 = orc.i32[_abi.I_EPISODE].copy()
    out = orc.step(np.zeros((4, 2), np.float32))
    assert out["done"].tolist() == [1, 1, 0, 0]
    assert out["opponents"].tolist() == [2, 1, 0, 0]
    assert (out["flags"][:2] & _abi.F_OPPONENT).all() and not (out["flags"][2:] & _abi.F_OPPONENT).any()
    assert out["time"][0] == pytest.approx(0.01) and out["time"][2] == pytest.approx(0.04)   # stopped at tick 1
    assert np.array_equal(orc.i32[_abi.I_EPISODE] - ep0, [1, 1, 0, 0])                        # whole world reset
    st = orc.read_stats()
    assert st["episodes"] == 1 and st["env_steps"] == 4 and st["collisions"] == 1


def test_one_done_car_ends_the_step_for_the_world_only():
    orc, tm = make(worlds=2, A=3, auto_reset=0, reset_mode=_abi.RESET_GRID, laps=10)
    orc.reset(mode=_abi.RESET_GRID)
    orc.f64[_abi.S_X, 4] += 1.0e3   # car 1 of world 1 far off the track: wall collision on the first tick
    out = orc.step(np.zeros((6, 2), np.float32))
    assert out["done"].tolist() == [0, 0, 0, 0, 1, 0]
    assert np.all((orc.i32[_abi.I_FLAGS, 3:] & _abi.F_NEEDS_RESET) != 0)      # the world waits for its reset ...
    assert not np.any(orc.i32[_abi.I_FLAGS, :3] & _abi.F_NEEDS_RESET)          # ... the other one carries on
    out = orc.step(np.zeros((6, 2), np.float32))
    assert out["done"].tolist() == [0, 0, 0, 1, 1, 1] and np.all(out["reward"][3:] == 0)


def test_time_limit_sets_every_done():
    orc, tm = make(worlds=1, A=2, time_limit_steps=3, auto_reset=0, reset_mode=_abi.RESET_GRID)
    orc.reset(mode=_abi.RESET_GRID)
    a = np.tile(np.array([[0.2, 0.0]], np.float32), (2, 1))
    assert orc.step(a)["done"].tolist() == [0, 0]
    assert orc.step(a)["done"].tolist() == [0, 0]
    assert orc.step(a)["done"].tolist() == [1, 1]
    assert orc.read_stats()["timeouts"] == 1


def test_rank_orders_the_world_by_lap_plus_progress():
    orc, tm = make(worlds=1, A=4, reset_mode=_abi.RESET_GRID, auto_reset=0)
    orc.reset(mode=_abi.RESET_GRID)
    out = orc.step(np.tile(np.array([[-1.0, 0.0]], np.float32), (4, 1)))    # nobody moves (motor 0.005)
    prog = orc.i32[_abi.I_LAP] + orc.f64[_abi.S_PROGRESS]
    order = np.lexsort((np.arange(4), -prog))
    want = np.empty(4, np.int32)
    want[order] = np.arange(1, 5)
    assert np.array_equal(out["rank"], want)
    assert out["rank"][3] == 1      # slot 3 stands furthest ahead on the grid


def test_n_step_progress_rewards_progress_over_n_ticks():
    T = _abi.TASK_N_STEP_PROGRESS
    orc, tm = make(worlds=1, A=2, tasks=[_abi.TASK_MAX_PROGRESS, T], reset_mode=_abi.RESET_GRID, auto_reset=0,
                   n_step_progress=10, action_repeat=4)
    orc.reset(mode=_abi.RESET_GRID)
    a = np.tile(np.array([[0.6, 0.0]], np.float32), (2, 1))
    P = [orc.i32[_abi.I_LAP, 1] + orc.f64[_abi.S_PROGRESS, 1]]
    want = []
    # replay the per-tick progress with single-tick steps of an identical env to know P_t
    ref, _ = make(worlds=1, A=2, tasks=[_abi.TASK_MAX_PROGRESS, T], reset_mode=_abi.RESET_GRID, auto_reset=0,
                  n_step_progress=10, action_repeat=1)
    ref.reset(mode=_abi.RESET_GRID)
    for t in range(120):
        ref.step(a)
        P.append(ref.i32[_abi.I_LAP, 1] + ref.f64[_abi.S_PROGRESS, 1])
    for k in range(30):
        before = orc.i32[_abi.I_LAP, 0] + orc.f64[_abi.S_PROGRESS, 0]
        out = orc.step(a)
        r = sum(100.0 * (P[t + 1] - P[max(t + 1 - 10, 0)]) for t in range(4 * k, 4 * k + 4))
        want.append(r)
        assert out["reward64"][1] == pytest.approx(r, rel=1e-9, abs=1e-12)
        # agent A keeps the one-tick maximize_progress reward: it telescopes to the progress made in this step
        after = orc.i32[_abi.I_LAP, 0] + orc.f64[_abi.S_PROGRESS, 0]
        assert out["reward64"][0] == pytest.approx(100.0 * (after - before), rel=1e-9, abs=1e-12)
    assert want[-1] > 0.0 and want[0] == 0.0   # progress is a per-cell quantity: nothing in the first 40 ms
    assert out["reward64"][1] > 2.0 * out["reward64"][0] > 0.0   # ten-tick windows overlap: about 10x the one-tick sum


def test_single_agent_path_is_untouched_by_the_world_fields():
    """agents_per_world = 1 and a non-n-step task: identical to the single-car oracle (whatever agent_task says)."""
    a = np.random.RandomState(0).uniform(-1, 1, (30, 16, 2)).astype(np.float32)
    outs = []
    for junk in (0, 1):
        orc, _ = make(worlds=16, A=1, reset_mode=_abi.RESET_RANDOM)
        if junk:
            for k in range(_abi.MAX_AGENTS):
                orc.cfg.agent_task[k] = _abi.TASK_N_STEP_PROGRESS
            orc.cfg.ball_spacing = 9.0
        orc.reset(mode=_abi.RESET_RANDOM)
        rec = []
        for t in range(30):
            o = orc.step(a[t])
            rec.append(np.concatenate([o["lidar"].ravel(), o["reward"], o["done"], o["progress"]]))
        outs.append(np.stack(rec))
    assert np.array_equal(outs[0], outs[1])


def test_fused_world_matches_the_reference_wrapper_stack(golden_dir):
    """The fused per-world step == the UNMODIFIED dict-of-agents wrappers of the reference (RaceCarWrapper, ActionRepeat
    `not any(dones.values())`, ReduceActionSpace, OccupancyMapObs, TimeLimit, Collect [REF dreamer/wrappers.py:55-250])
    run tick by tick over the one-tick multi-car world (tests/golden/make_golden.py::multi_agent_stack_golden)."""
    import helpers
    g = np.load(golden_dir / "multi_agent_stack_golden.npz")
    cfg = helpers.fused_world_config(default_config(), g)
    orc = Oracle(cfg, [load_track("austria")])
    rec = helpers.replay_world(lambda: orc.reset(mode=_abi.RESET_RANDOM_BALL), lambda a: orc.step(a), g["actions"],
                               g["reset_before"])
    helpers.assert_matches_multi_agent_golden(rec, g)
    assert (g["opponents"] != 0).sum() > 20 and g["done"].all(1).sum() >= 3   # contacts and time-outs are in the fixture


def test_world_invariants_under_random_play():
    """Laws of the world step that need no reference: contact is mutual, ranks are a permutation of 1..A per world, a car
    never reports contact with itself, the cars of a world share time, step counter and episode counter."""
    A, worlds = 4, 96
    orc, tm = make(worlds=worlds, A=A, ball_spacing=0.7, time_limit_steps=25, seed=11)
    orc.reset(mode=_abi.RESET_RANDOM_BALL)
    rng = np.random.RandomState(3)
    gain = np.tile(np.array([1.0, 0.15, 0.6, 0.05], np.float32), worlds)
    contacts = 0
    for k in range(60):
        a = rng.uniform(-1, 1, (A * worlds, 2)).astype(np.float32)
        a[:, 0] = ((np.abs(a[:, 0]) * 0.5 + 0.5) * gain) * 2 - 1
        a[:, 1] *= 0.35
        out = orc.step(a)
        opp = out["opponents"].reshape(worlds, A)
        for i in range(A):
            assert not np.any(opp[:, i] >> i & 1)                               # never with itself
            for j in range(A):
                assert np.array_equal(opp[:, i] >> j & 1, opp[:, j] >> i & 1)   # mutual
        assert np.array_equal(np.sort(out["rank"].reshape(worlds, A), axis=1), np.tile(np.arange(1, A + 1), (worlds, 1)))
        assert np.array_equal((out["flags"] & _abi.F_OPPONENT) != 0, out["opponents"] != 0)
        for row in (orc.f64[_abi.S_TIME], orc.i32[_abi.I_AGENT_STEP], orc.i32[_abi.I_EPISODE]):
            w = row.reshape(worlds, A)
            assert np.all(w == w[:, :1])
        contacts += int((opp != 0).sum())
    assert contacts > 50
