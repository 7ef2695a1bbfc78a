"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI (librd_env.so) and is
compared with the CPU oracle on the same seeded inputs, and with the golden fixtures recorded from the reference.

Tolerances (BASELINE.json north_star): occupancy grids and termination/lap/wrong-way flags bit-exact; LiDAR ranges
within 1e-3 m; dynamics state within 1e-5 relative after N steps.
"""
import os

import numpy as np
import pytest

import helpers
from racing_dreamer_b200 import _abi

pytestmark = pytest.mark.gpu

LIDAR_TOL_M = 1e-3
DYN_RTOL = 1e-5
THREADS = os.cpu_count() or 1


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


def make_env(torch, **kw):
    from racing_dreamer_b200 import BatchedRaceEnv, EnvConfig
    return BatchedRaceEnv(EnvConfig(**kw), device="cuda:0")


def make_oracle(env, n_threads=THREADS):
    from oracle import Oracle
    return Oracle(env.cfg, env.tracks, env.map_ids, n_threads=n_threads)


def random_poses(tm, n, rng, jitter=0.15, any_yaw=True):
    p = tm.reset_poses[rng.randint(0, len(tm.reset_poses), n)].copy()
    p[:, :2] += rng.uniform(-jitter, jitter, (n, 2))
    p[:, 2] = rng.uniform(-np.pi, np.pi, n) if any_yaw else p[:, 2] + rng.uniform(-0.6, 0.6, n)
    return p


# ---------------------------------------------------------------------------------------------- a2 LiDAR
@pytest.mark.parametrize("track", ["austria", "columbia", "treitlstrasse_v2", "barcelona", "gbr"])
def test_lidar_vs_oracle(torch_cuda, track):
    torch = torch_cuda
    env = make_env(torch, tracks=(track,), n_envs=8)
    orc = make_oracle(env)
    tm = env.tracks[0]
    rng = np.random.RandomState(11)
    poses = random_poses(tm, 1024, rng)
    poses[0] = tm.start_poses[0]
    poses[1] = (1e6, 1e6, 0.3)                 # far outside the map
    poses[2] = (poses[2, 0], poses[2, 1], 0.0)  # axis-aligned beams
    poses[3] = (poses[3, 0], poses[3, 1], np.pi / 2)
    got = env.lidar_cast(torch.from_numpy(poses)).cpu().numpy()
    want = orc.lidar_cast(poses)
    assert got.shape == (1024, 1080)
    assert np.abs(got - want).max() <= LIDAR_TOL_M
    assert (got != want).mean() < 1e-4          # in practice bit-identical: the march is integer-only
    assert got.min() >= 0.25 - 1e-6 and got.max() <= 15.0
    assert np.all(got[1] == np.float32(0.25))   # sensor outside the drivable area: every beam reads range_min
    env.close()


def test_lidar_mixed_maps_ragged_and_empty(torch_cuda):
    torch = torch_cuda
    env = make_env(torch, tracks=("barcelona", "austria"), n_envs=4)
    orc = make_oracle(env)
    rng = np.random.RandomState(5)
    pa = random_poses(env.tracks[0], 37, rng)      # ragged: not a multiple of anything
    pb = random_poses(env.tracks[1], 91, rng)
    poses = np.concatenate([pa, pb])
    ids = np.array([0] * 37 + [1] * 91, np.int32)
    got = env.lidar_cast(torch.from_numpy(poses), ids).cpu().numpy()
    assert np.abs(got - orc.lidar_cast(poses, ids)).max() <= LIDAR_TOL_M
    assert env.lidar_cast(torch.zeros((0, 3), dtype=torch.float64)).shape == (0, 1080)   # empty input
    with pytest.raises(RuntimeError):
        env.lidar_cast(torch.from_numpy(poses), ids[::-1].copy())                          # ids must be ascending
    env.close()


def test_lidar_properties_full_size(torch_cuda):
    """Size-independent properties at BASELINE config-2 size (4096 envs x 1080 beams): range bounds; turning the car
    by exactly one beam step shifts the scan by one index (up to re-quantisation of the ray); idempotence."""
    torch = torch_cuda
    env = make_env(torch, tracks=("austria",), n_envs=8)
    tm = env.tracks[0]
    rng = np.random.RandomState(2)
    poses = random_poses(tm, 4096, rng)
    a = env.lidar_cast(torch.from_numpy(poses)).cpu().numpy()
    assert a.min() >= np.float32(0.25) and a.max() <= np.float32(15.0)
    step = float(env.cfg.lidar_fov) / 1079.0
    shifted = poses.copy()
    shifted[:, 2] -= step                      # heading turned right by one beam: beam i now looks where beam i+1 did
    b = env.lidar_cast(torch.from_numpy(shifted)).cpu().numpy()
    d = np.abs(b[:, :-1] - a[:, 1:])
    assert np.median(d) < 1e-3 and (d < 0.05).mean() > 0.97
    assert np.array_equal(a, env.lidar_cast(torch.from_numpy(poses)).cpu().numpy())
    env.close()


def test_lidar_noise_and_normalisation(torch_cuda):
    torch = torch_cuda
    env = make_env(torch, tracks=("columbia",), n_envs=8, lidar_noise=0.03, normalize_lidar=True, seed=99)
    orc = make_oracle(env)
    poses = random_poses(env.tracks[0], 256, np.random.RandomState(8))
    got = env.lidar_cast(torch.from_numpy(poses)).cpu().numpy()
    want = orc.lidar_cast(poses)
    assert np.abs(got - want).max() <= LIDAR_TOL_M / 15.0
    assert got.min() >= -0.5 and got.max() <= 0.5       # r/15 - 0.5 [REF dreamer/tools.py:274]
    env.close()


# ---------------------------------------------------------------------------------------------- a1 dynamics
def test_dynamics_vs_oracle(torch_cuda):
    torch = torch_cuda
    env = make_env(torch, tracks=("austria",), n_envs=8)
    orc = make_oracle(env)
    rng = np.random.RandomState(4)
    n = 2048
    state = np.zeros((7, n))
    state[0] = rng.uniform(-5, 5, n)
    state[1] = rng.uniform(-5, 5, n)
    state[2] = rng.uniform(-0.4, 0.4, n)
    state[3] = rng.uniform(0.0, 4.5, n)
    state[4] = rng.uniform(-np.pi, np.pi, n)
    state[5] = rng.uniform(-1, 1, n)
    state[6] = rng.uniform(-0.2, 0.2, n)
    cmd = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)], 1)
    for ticks in (1, 8, 400):
        got = env.dynamics(torch.from_numpy(state), torch.from_numpy(cmd), ticks).cpu().numpy()
        want = orc.dynamics(state, cmd, ticks)
        scale = np.maximum(np.abs(want), 1.0)
        assert np.all(np.abs(got - want) <= DYN_RTOL * scale), f"ticks={ticks}: {np.abs(got - want).max()}"
        assert np.abs(got - want).max() < 1e-9          # in practice ~1e-13: same equations, same op order
    env.close()


# ---------------------------------------------------------------------------------------------- a5 occupancy
def test_occupancy_vs_golden_and_oracle(torch_cuda, golden_dir):
    torch = torch_cuda
    g = np.load(golden_dir / "occupancy_golden.npz")
    names = [str(n) for n in g["track_names"]]
    env = make_env(torch, tracks=tuple(names), n_envs=8, obs_type="lidar_occupancy")
    got = env.occupancy_obs(torch.from_numpy(g["poses"]), g["track"]).cpu().numpy()
    want = np.unpackbits(g["images"], axis=2)[:, :, :64]
    assert np.array_equal(got, want), f"{(got != want).sum()} px differ from the reference's OccupancyMapObs"
    orc = make_oracle(env)
    rng = np.random.RandomState(21)
    poses = np.concatenate([random_poses(t, 100, rng) for t in env.tracks])
    ids = np.repeat(np.arange(3, dtype=np.int32), 100)
    got = env.occupancy_obs(torch.from_numpy(poses), ids).cpu().numpy()
    assert np.array_equal(got, orc.occupancy_obs(poses, ids))
    env.close()


def test_occupancy_exact_path_forced(torch_cuda, golden_dir, monkeypatch):
    """k_occupancy evaluates pixels in float32 and re-evaluates, in float64, the ones that land within eps of a rounding
    threshold.  Widening eps to 0.05 (RD_OCC_EPS) pushes every edge pixel through that float64 path: the images must
    still be the reference's, bit for bit -- this pins the exact evaluator, which the default eps exercises only about
    once in 40 images."""
    torch = torch_cuda
    monkeypatch.setenv("RD_OCC_EPS", "0.05")
    g = np.load(golden_dir / "occupancy_golden.npz")
    names = [str(n) for n in g["track_names"]]
    env = make_env(torch, tracks=tuple(names), n_envs=8, obs_type="lidar_occupancy")
    got = env.occupancy_obs(torch.from_numpy(g["poses"]), g["track"]).cpu().numpy()
    want = np.unpackbits(g["images"], axis=2)[:, :, :64]
    assert np.array_equal(got, want), f"{(got != want).sum()} px differ from the reference's OccupancyMapObs"
    env.close()


def test_occupancy_many_poses_vs_oracle(torch_cuda):
    """4096 random poses per track (incl. map-edge crops and the four axis-aligned headings) against the float64 oracle."""
    torch = torch_cuda
    env = make_env(torch, tracks=("austria", "treitlstrasse_v2"), n_envs=8, obs_type="lidar_occupancy")
    orc = make_oracle(env)
    rng = np.random.RandomState(33)
    poses = np.concatenate([random_poses(t, 2048, rng, jitter=0.3) for t in env.tracks])
    poses[:4, 2] = (0.0, np.pi / 2, np.pi, -np.pi / 2)
    poses[4] = (1e4, 1e4, 0.1)                       # crop entirely outside the map -> all zeros
    ids = np.repeat(np.arange(2, dtype=np.int32), 2048)
    got = env.occupancy_obs(torch.from_numpy(poses), ids).cpu().numpy()
    want = orc.occupancy_obs(poses, ids)
    assert np.array_equal(got, want), f"{(got != want).any(axis=(1, 2)).sum()} of {len(poses)} images differ"
    assert not got[4].any()
    env.close()


# ---------------------------------------------------------------------------------------------- fused step
def _gpu_step_fn(torch, env):
    def step(a):
        obs, rew, done, info = env.step(torch.from_numpy(np.ascontiguousarray(a)).cuda())
        out = {"lidar": obs["lidar"], "pose": obs["pose"], "velocity": obs["velocity"], "speed": obs["speed"],
               "reward": rew, "done": done.to(torch.uint8), "progress": info["progress"], "lap": info["lap"],
               "time": info["time"], "flags": info["flags"]}
        if "lidar_occupancy" in obs:
            out["occupancy"] = obs["lidar_occupancy"][..., 0]
        return {k: v.cpu().numpy() for k, v in out.items()}
    return step


def test_dreamer_stack_golden_replay(torch_cuda, golden_dir):
    """BASELINE config 1 through the CUDA path == the reference wrapper stack, step for step."""
    torch = torch_cuda
    from racing_dreamer_b200 import BatchedRaceEnv
    g = np.load(golden_dir / "dreamer_stack_golden.npz")
    cfg = helpers.fused_dreamer_config(_abi.default_config(), int(g["action_repeat"]), int(g["duration"]))
    env = BatchedRaceEnv(tracks=("columbia",), raw_config=cfg, device="cuda:0")
    rec = helpers.replay(lambda: env.reset(mode="grid"), _gpu_step_fn(torch, env), g["actions"], g["reset_before"])
    helpers.assert_matches_dreamer_golden(rec, g, lidar_tol=LIDAR_TOL_M, float_tol=DYN_RTOL)
    env.close()


def test_baselines_stack_golden_replay(torch_cuda, golden_dir):
    torch = torch_cuda
    from racing_dreamer_b200 import BatchedRaceEnv
    g = np.load(golden_dir / "baselines_stack_golden.npz")
    cfg = helpers.fused_baselines_config(_abi.default_config(), int(g["repeat"]))
    env = BatchedRaceEnv(tracks=("austria",), raw_config=cfg, device="cuda:0")
    rec = helpers.replay(lambda: env.reset(mode="grid"), _gpu_step_fn(torch, env), g["actions"], g["reset_before"])
    helpers.assert_matches_baselines_golden(rec, g, float_tol=DYN_RTOL, lidar_tol=LIDAR_TOL_M)
    env.close()


def _compare_state(env, orc, step_idx):
    f, i = env.get_state()
    f, i = f.cpu().numpy(), i.cpu().numpy()
    assert np.array_equal(i, orc.i32), f"integer state (lap/checkpoint/flags/...) differs at step {step_idx}"
    scale = np.maximum(np.abs(orc.f64), 1.0)
    assert np.all(np.abs(f - orc.f64) <= DYN_RTOL * scale), f"float64 state differs at step {step_idx}"


@pytest.mark.parametrize("track,n,repeat,mode", [("treitlstrasse_v2", 4096, 8, "random"), ("austria", 1024, 4, "grid")])
def test_closed_loop_vs_oracle_with_auto_reset(torch_cuda, track, n, repeat, mode):
    """BASELINE config 4 subsample: random actions so that collisions / time limits fire; auto-reset on; the whole
    integer state (lap, checkpoint, flags, counters) must stay bit-identical, flags/done bit-exact every step."""
    torch = torch_cuda
    env = make_env(torch, tracks=(track,), n_envs=n, action_repeat=repeat, auto_reset=True, reset_mode=mode,
                   time_limit_steps=40, seed=4, laps=1)
    orc = make_oracle(env)
    env.reset()
    orc.reset(mode=int(env.cfg.reset_mode))
    _compare_state(env, orc, -1)
    rng = np.random.RandomState(4)
    dones = 0
    for k in range(60):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        a[:, 0] = np.abs(a[:, 0])
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"])
        assert np.array_equal(info["flags"].cpu().numpy(), ref["flags"])
        assert np.array_equal(info["lap"].cpu().numpy(), ref["lap"])
        assert np.allclose(rew.cpu().numpy(), ref["reward"], rtol=DYN_RTOL, atol=1e-6)
        assert np.abs(obs["lidar"].cpu().numpy() - ref["lidar"]).max() <= LIDAR_TOL_M
        assert np.allclose(obs["speed"].cpu().numpy(), ref["speed"], rtol=DYN_RTOL, atol=1e-7)
        _compare_state(env, orc, k)
        dones += int(ref["done"].sum())
    assert dones > n // 4, "the scenario is supposed to exercise terminations"
    gs, os_ = env.read_stats(), orc.stats.as_dict()
    for key in gs:
        assert abs(gs[key] - os_[key]) <= 1e-6 * max(1.0, abs(os_[key])), key
    env.close()


def test_mixed_maps_step_and_occupancy(torch_cuda):
    """BASELINE config 5 layout: Barcelona/Austria alternating by env index; plus the occupancy obs in the step."""
    torch = torch_cuda
    n = 96
    env = make_env(torch, tracks=("barcelona", "austria"), n_envs=n, action_repeat=8, obs_type="lidar_occupancy",
                   auto_reset=True, reset_mode="random", seed=3)
    assert np.array_equal(env.map_ids, np.arange(n) % 2)
    orc = make_oracle(env)
    o = env.reset()
    r = orc.reset(mode=int(env.cfg.reset_mode))
    assert not o["lidar_occupancy"].any()                        # reset obs is zeros [REF dreamer/wrappers.py:410-414]
    assert np.abs(o["lidar"].cpu().numpy() - r["lidar"]).max() <= LIDAR_TOL_M
    rng = np.random.RandomState(9)
    for k in range(12):
        a = np.stack([rng.uniform(0.2, 1.0, n), rng.uniform(-0.5, 0.5, n)], 1).astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert np.array_equal(obs["lidar_occupancy"].cpu().numpy()[..., 0], ref["occupancy"])
        assert np.abs(obs["lidar"].cpu().numpy() - ref["lidar"]).max() <= LIDAR_TOL_M
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"])
    env.close()


def test_manual_reset_mask_frozen_envs_and_state_roundtrip(torch_cuda):
    torch = torch_cuda
    n = 64
    env = make_env(torch, tracks=("columbia",), n_envs=n, action_repeat=4, auto_reset=False, time_limit_steps=5)
    orc = make_oracle(env)
    env.reset(mode="random")
    orc.reset(mode=_abi.RESET_RANDOM)
    a = np.tile(np.array([[0.5, 0.1]], np.float32), (n, 1))
    for k in range(7):                                           # TimeLimit fires at step 5, then envs are frozen
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"])
        assert np.allclose(rew.cpu().numpy(), ref["reward"], rtol=DYN_RTOL, atol=1e-6)
    assert done.all()
    mask = np.zeros(n, np.uint8)
    mask[::3] = 1
    env.reset(mask=torch.from_numpy(mask).cuda(), mode="grid")
    orc.reset(mask=mask, mode=_abi.RESET_GRID)
    _compare_state(env, orc, 100)
    obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
    ref = orc.step(a)
    assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"]) and done.cpu().numpy()[1]
    f, i = env.get_state()
    env2 = make_env(torch, tracks=("columbia",), n_envs=n, action_repeat=4, auto_reset=False, time_limit_steps=5)
    env2.set_state(f, i)                                          # checkpoint / resume
    o1 = env.step(torch.from_numpy(a).cuda())
    r1 = o1[1].clone()
    o2 = env2.step(torch.from_numpy(a).cuda())
    assert torch.equal(r1, o2[1]) and torch.equal(env.get_state()[0], env2.get_state()[0])
    env.close()
    env2.close()


def test_step_before_reset_is_an_error(torch_cuda):
    torch = torch_cuda
    env = make_env(torch, tracks=("austria",), n_envs=4)
    with pytest.raises(RuntimeError, match="Must reset environment"):   # [REF dreamer/wrappers.py:148]
        env.step(torch.zeros((4, 2), device="cuda"))
    env.close()


def test_max_speed_task(torch_cuda):
    """In-tree task MaximizeSpeed [REF baselines/racing/environment/tasks.py:4-22]: never done, -exp(|steer| - v_x)."""
    torch = torch_cuda
    n = 128
    env = make_env(torch, tracks=("austria",), n_envs=n, action_repeat=1, task="max_speed", auto_reset=False,
                   rescale_actions=False)
    orc = make_oracle(env)
    env.reset()
    orc.reset()
    rng = np.random.RandomState(1)
    for k in range(30):
        a = np.stack([rng.uniform(0.0, 1.0, n), rng.uniform(-0.3, 0.3, n)], 1).astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert not done.any()
        assert np.allclose(rew.cpu().numpy(), ref["reward"], rtol=1e-5, atol=1e-6)
    v = obs["velocity"].cpu().numpy()[:, 0].astype(np.float64)
    col = info["wall_collision"].cpu().numpy()
    want = np.where(col, -1.0, -np.exp(np.abs(a[:, 1].astype(np.float64)) - v))
    assert np.allclose(rew.cpu().numpy(), want, rtol=1e-4, atol=1e-5)
    env.close()


def test_float16_scans_are_the_rounded_float32_scans(torch_cuda):
    """lidar_dtype='float16' = Collect._convert at precision 16 [REF dreamer/wrappers.py:240-250; dream.py:176-177]: every
    range is the float32 value rounded to nearest-even half, on the device path, the stage entry and the host-facing path;
    the on-device policies (float32 readers) refuse such an env."""
    torch = torch_cuda
    from racing_dreamer_b200 import EnvConfig, GapFollowerPolicy
    from racing_dreamer_b200.host import HostSteppedEnv
    n = 300
    kw = dict(tracks=("austria", "columbia"), n_envs=n, action_repeat=4, auto_reset=True, reset_mode="random", seed=8,
              time_limit_steps=20, lidar_noise=0.03)
    e32, e16 = make_env(torch, **kw), make_env(torch, lidar_dtype="float16", **kw)
    h16 = HostSteppedEnv(EnvConfig(lidar_dtype="float16", **kw), device="cuda:0", n_shards=3)
    orc = make_oracle(e32)
    o32, o16 = e32.reset(), e16.reset()
    ho = h16.reset()
    ref = orc.reset(mode=int(e32.cfg.reset_mode))
    assert o16["lidar"].dtype == torch.float16 and ho["lidar"].dtype == np.float16
    rng = np.random.RandomState(1)
    for k in range(25):
        assert torch.equal(o16["lidar"], o32["lidar"].half()), k
        assert np.array_equal(ho["lidar"], o16["lidar"].cpu().numpy()), k
        assert np.array_equal(ho["lidar"], ref["lidar"].astype(np.float16)), k          # numpy's cast is the specification
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        o32, o16 = e32.step(torch.from_numpy(a).cuda())[0], e16.step(torch.from_numpy(a).cuda())[0]
        ho = h16.step(a)
        ref = orc.step(a)
    poses = random_poses(e32.tracks[0], 64, rng)
    assert torch.equal(e16.lidar_cast(torch.from_numpy(poses)), e32.lidar_cast(torch.from_numpy(poses)).half())
    assert h16.d2h_bytes_per_step < n * 1080 * 2 + n * 200
    with pytest.raises(RuntimeError, match="float32 scans"):
        GapFollowerPolicy(e16)
    e32.close(); e16.close(); h16.close()


def test_lidar_kernel_variants_mix_on_one_handle(torch_cuda):
    """One handle launches different k_lidar instantiations depending on the launch size (two beam groups per work item for
    long launches) --
    e.g. the full batch in rd_step and small chunks in rd_step_host, or a large and a small rd_lidar_cast.  Every
    instantiation must be opted in to the large shared-memory carve-out on its own (regression: 'invalid argument')."""
    torch = torch_cuda
    from racing_dreamer_b200 import EnvConfig
    from racing_dreamer_b200.host import HostSteppedEnv
    rng = np.random.RandomState(4)
    for track in ("barcelona", "austria", "treitlstrasse_v2"):
        env = make_env(torch, tracks=(track,), n_envs=8)
        orc = make_oracle(env)
        big = random_poses(env.tracks[0], 40000, rng)              # >= 45 beam groups per resident warp: two groups per item
        small = big[:64]
        a = env.lidar_cast(torch.from_numpy(small)).cpu().numpy()   # plain variant first ...
        b = env.lidar_cast(torch.from_numpy(big)).cpu().numpy()     # ... then the other one on the same handle
        c = env.lidar_cast(torch.from_numpy(small)).cpu().numpy()
        assert np.array_equal(a, b[:64]) and np.array_equal(a, c)
        assert np.abs(b[:2048] - orc.lidar_cast(big[:2048])).max() <= LIDAR_TOL_M
        env.close()
    # the host-facing path: k_step over 24576 envs, scans in chunks of very different sizes
    ec = EnvConfig(tracks=("barcelona", "austria"), n_envs=24576, action_repeat=4, auto_reset=True, reset_mode="random", seed=2)
    dev, host = make_env(torch, **{k: getattr(ec, k) for k in ("tracks", "n_envs", "action_repeat", "auto_reset", "reset_mode", "seed")}), \
        HostSteppedEnv(ec, device="cuda:0", n_shards=8)
    dev.reset(); host.reset()
    for k in range(3):
        a = rng.uniform(-1, 1, (24576, 2)).astype(np.float32)
        obs = dev.step(torch.from_numpy(a).cuda())[0]
        out = host.step(a)
        assert np.array_equal(out["lidar"], obs["lidar"].cpu().numpy()), k
    dev.close(); host.close()


@pytest.mark.parametrize("nb", [1, 31, 32, 33, 100, 1081])
def test_lidar_ragged_beam_counts(torch_cuda, nb):
    """Beam counts that are not multiples of the 32-beam work item (one group, a ragged last group, odd group counts in the
    longest-first order) and the degenerate single beam, through the step and the stage entry."""
    torch = torch_cuda
    n = 160
    env = make_env(torch, tracks=("treitlstrasse_v2", "austria"), n_envs=n, n_beams=nb, action_repeat=2, auto_reset=True,
                   reset_mode="random", seed=nb, time_limit_steps=6)
    orc = make_oracle(env)
    o, r = env.reset(), orc.reset(mode=int(env.cfg.reset_mode))
    assert o["lidar"].shape == (n, nb) and np.abs(o["lidar"].cpu().numpy() - r["lidar"]).max() <= LIDAR_TOL_M
    rng = np.random.RandomState(nb)
    for k in range(8):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert np.abs(obs["lidar"].cpu().numpy() - ref["lidar"]).max() <= LIDAR_TOL_M
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"])
    poses = random_poses(env.tracks[1], 300, rng)
    ids = np.ones(300, np.int32)
    assert np.abs(env.lidar_cast(torch.from_numpy(poses), ids).cpu().numpy() - orc.lidar_cast(poses, ids)).max() <= LIDAR_TOL_M
    env.close()


# ---------------------------------------------------------------------------------------------- round-2 additions
def test_dynamics_vs_independent_numpy_model(torch_cuda):
    """north_star: "dynamics state after N steps must agree within 1e-5 relative, compared against a float64 NumPy
    rendition of the same model" -- tests/np_single_track.py, the textbook form of SURVEY.md Appendix C that shares no
    code or operation order with the kernel or the C oracle."""
    from np_single_track import integrate, params_from_config, random_states
    torch = torch_cuda
    env = make_env(torch, tracks=("austria",), n_envs=8)
    p = params_from_config(env.cfg)
    for seed, (v_lo, v_hi), ticks in ((4, (0.0, 4.5), 400), (12, (0.3, 0.7), 60), (13, (0.0, 4.5), 8)):
        state, cmd = random_states(4096, np.random.RandomState(seed), v_lo=v_lo, v_hi=v_hi)
        got = env.dynamics(torch.from_numpy(state), torch.from_numpy(cmd), ticks).cpu().numpy()
        want = integrate(p, state, cmd, ticks, dt=float(env.cfg.dt))
        err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
        assert err.max() <= DYN_RTOL, f"seed {seed}: {err.max():.3e}"      # the contract
        assert err.max() < 1e-9, f"seed {seed}: {err.max():.3e}"           # in practice ~1e-13
    env.close()


def test_reward_done_stage_vs_oracle(torch_cuda):
    """a7/a8 stage entry rd_reward_done: teacher-forced poses along and across the track, several ticks in a row so that
    checkpoints advance, laps complete, cars leave the track; integer bookkeeping bit-exact, reward to rounding."""
    torch = torch_cuda
    for task, kw in (("maximize_progress", dict(laps=1, n_checkpoints=20)), ("max_speed", {}),
                     ("maximize_progress", dict(progress_abs=True, terminate_on_collision=False, n_checkpoints=4, time_limit=0.05))):
        env = make_env(torch, tracks=("treitlstrasse_v2", "austria"), n_envs=8, task=task, **kw)
        orc = make_oracle(env)
        rng = np.random.RandomState(31)
        n = 3000
        ids = np.sort(rng.randint(0, 2, n)).astype(np.int32)
        bf = np.zeros((3, n)); bi = np.zeros((3, n), np.int32)
        bi[0] = 1
        idx = [rng.randint(0, len(env.tracks[m].reset_poses)) for m in ids]
        bf[1] = rng.uniform(0, 1, n); bf[2] = 1.0 + bf[1]; bi[1] = (bf[1] * int(env.cfg.n_checkpoints)).astype(np.int32)
        gbf, gbi = torch.from_numpy(bf), torch.from_numpy(bi)
        for k in range(6):
            stride = 40 * k                                   # walk along the lap: checkpoints and the finish line pass by
            poses = np.stack([env.tracks[m].reset_poses[(i + stride) % len(env.tracks[m].reset_poses)] for m, i in zip(ids, idx)])
            kin = np.zeros((5, n))
            kin[0] = poses[:, 0] + rng.uniform(-0.4, 0.4, n); kin[1] = poses[:, 1] + rng.uniform(-0.4, 0.4, n)
            kin[2] = poses[:, 2] + rng.uniform(-1, 1, n); kin[3] = rng.uniform(0, 4, n); kin[4] = rng.uniform(-0.3, 0.3, n)
            kin[0, :5] = 1e7                                  # far outside the map
            steer = rng.uniform(-1, 1, n)
            gbf, gbi, grew, gdone = env.reward_done(torch.from_numpy(kin), torch.from_numpy(steer), gbf, gbi, ids)
            bf, bi, rew, done = orc.reward_done(kin, steer, bf, bi, ids)
            assert np.array_equal(gbi.cpu().numpy(), bi), (task, k)
            assert np.array_equal(gdone.cpu().numpy(), done), (task, k)
            assert np.array_equal(gbf.cpu().numpy(), bf), (task, k)          # time, progress, lap + progress: same operations
            assert np.allclose(grew.cpu().numpy(), rew, rtol=1e-12, atol=1e-12), (task, k)
        assert bi[0].max() >= 2 and (bi[2] & _abi.F_COLLISION).any() and (bi[2] & _abi.F_WRONG_WAY).any()
        env.close()


def test_baselines_normalize_observations(torch_cuda):
    """RD_OBS_NORM_BASELINES: lidar / pose / velocity = (x - low) * (1 / (high - low)) in float64, then float32
    [REF baselines/racing/environment/single_agent.py:66-99], fused into the stores of k_lidar and k_step."""
    torch = torch_cuda
    n = 256
    kw = dict(tracks=("austria",), n_envs=n, action_repeat=4, auto_reset=True, reset_mode="random", seed=5,
              clip_actions=True, rescale_actions=False, repeat_semantics="baselines", time_limit_ticks=50)
    env = make_env(torch, normalize_obs="baselines", **kw)
    raw = make_env(torch, **kw)
    orc = make_oracle(env)
    env.reset(); raw.reset(); orc.reset(mode=_abi.RESET_RANDOM)
    rng = np.random.RandomState(3)
    for k in range(20):
        a = rng.uniform(-1.2, 1.2, (n, 2)).astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        robs, _, rdone, _ = raw.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"])
        assert torch.equal(done, rdone)
        for key in ("lidar", "pose", "velocity"):
            got = obs[key].cpu().numpy()
            assert np.abs(got - ref[key]).max() <= 1e-6, key
            assert got.min() >= 0.0 and got.max() <= 1.0, key
        # against the un-normalised run of the same kernels: exactly NormalizeObservations' arithmetic
        want = ((robs["lidar"].cpu().numpy().astype(np.float64) - 0.0) * (1.0 / (15.0 - 0.0))).astype(np.float32)
        assert np.array_equal(obs["lidar"].cpu().numpy(), want)
    st = env.read_stats()
    assert st["timeouts"] > 0                              # gym TimeLimit(50 ticks) inside ActionRepeat fired
    env.close(); raw.close()


def test_public_step_launches_only_the_env_kernels(torch_cuda):
    """BatchedRaceEnv.step() -- the public API -- must not run eager PyTorch on the hot path: every tensor it returns is a
    persistent buffer the CUDA kernels wrote (done and the two info flags as bool views of 0/1 bytes)."""
    torch = torch_cuda
    from torch.profiler import profile, ProfilerActivity
    env = make_env(torch, tracks=("austria",), n_envs=512, action_repeat=8, auto_reset=True, reset_mode="random")
    env.reset()
    a = torch.zeros((512, 2), device="cuda")
    env.step(a)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            obs, rew, done, info = env.step(a)
        torch.cuda.synchronize()
    names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    kernels = [nm for nm in names if "memcpy" not in nm.lower() and "memset" not in nm.lower()]
    assert kernels, "the profiler saw no kernels"
    assert all(nm.startswith(("k_step", "k_lidar", "void k_lidar", "k_occupancy")) for nm in kernels), sorted(set(kernels))
    assert done.dtype == torch.bool and done.data_ptr() == env.buf["done"].data_ptr()
    assert info["wrong_way"].data_ptr() == env.buf["wrong_way"].data_ptr()
    env.close()


@pytest.mark.parametrize("name", ["train", "test"])
def test_baselines_chain_golden_replay(torch_cuda, golden_dir, name):
    """The model-free wrap chains of the reference (recorded with its own classes, baselines_chain_golden.npz) through
    the CUDA path: clip, NormalizeObservations, tick-based gym TimeLimit inside ActionRepeat, the baselines repeat rule."""
    torch = torch_cuda
    from racing_dreamer_b200 import BatchedRaceEnv
    g = np.load(golden_dir / "baselines_chain_golden.npz")
    cfg = helpers.fused_baselines_chain_config(_abi.default_config(), g, test=(name == "test"))
    env = BatchedRaceEnv(tracks=("treitlstrasse_v2",), raw_config=cfg, device="cuda:0")
    rec = helpers.replay(lambda: env.reset(mode="grid"), _gpu_step_fn(torch, env), g["actions"], g[f"{name}_reset_before"])
    helpers.assert_matches_baselines_chain_golden(rec, g, name, lidar_tol=LIDAR_TOL_M / 15.0, float_tol=DYN_RTOL)
    env.close()


def test_simulate_statistics_on_device(torch_cuda, golden_dir):
    """a12: rd_read_stats reproduces what the reference's tools.simulate collects per episode -- max over the episode of
    lap + progress - 1 and the cumulative reward [REF dreamer/tools.py:178-199] (simulate_golden.npz, recorded from the
    unmodified function)."""
    torch = torch_cuda
    from racing_dreamer_b200 import BatchedRaceEnv
    g = np.load(golden_dir / "simulate_golden.npz")
    cfg = helpers.fused_dreamer_config(_abi.default_config(), int(g["action_repeat"]), int(g["duration"]), occupancy=False)
    env = BatchedRaceEnv(tracks=("treitlstrasse_v2",), raw_config=cfg, device="cuda:0")
    returns, maxima = helpers.simulate_statistics(lambda: env.reset(mode="grid"), _gpu_step_fn(torch, env),
                                                  lambda: env.read_stats(reset=True), g)
    assert np.allclose(returns, g["cum_rewards"], rtol=1e-9, atol=1e-9)
    assert np.allclose(maxima, g["max_progresses"], rtol=0, atol=1e-9)
    env.close()


@pytest.mark.parametrize("kw", [dict(auto_reset=True, reset_mode="random"), dict(auto_reset=False, reset_mode="grid"),
                                dict(auto_reset=True, reset_mode="random_bidirectional", repeat_semantics="baselines",
                                     time_limit_ticks=96, normalize="baselines")])
def test_split_step_kernel_is_bitwise_the_single_warp_kernel(torch_cuda, monkeypatch, kw):
    """k_step_split (five warps per 32 envs: dynamics | position | probe | bookkeeping | reset look-ahead) and k_step
    (one warp) are the same arithmetic in a different instruction order: every output, the whole env state and the
    episode statistics must agree bit for bit, step after step, through terminations, auto-resets, frozen envs and a
    ragged last warp."""
    torch = torch_cuda
    kw = dict(kw)
    if kw.pop("normalize", None) == "baselines":
        kw["normalize_obs"] = "baselines"
    n = 1000 + 13                                          # ragged: the last CTA has 21 live lanes
    envs = []
    for split in ("1", "0"):
        monkeypatch.setenv("RD_STEP_SPLIT", split)
        envs.append(make_env(torch, tracks=("austria",), n_envs=n, action_repeat=8, time_limit_steps=25, seed=9, laps=1, **kw))
    for e in envs:
        e.reset()
    rng = np.random.RandomState(21)
    finished = 0
    for k in range(40):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        a[:, 0] = np.abs(a[:, 0])
        outs = []
        for e in envs:
            obs, rew, done, info = e.step(torch.from_numpy(a).cuda())
            f64, i32 = e.get_state()
            outs.append(([obs[key].cpu().numpy() for key in sorted(obs)], rew.cpu().numpy(), done.cpu().numpy(),
                         [info[key].cpu().numpy() for key in sorted(info)], f64.cpu().numpy(), i32.cpu().numpy()))
        x, y = outs
        for u, v in zip(x[0], y[0]):
            assert np.array_equal(u, v), f"step {k}: observation differs"
        assert np.array_equal(x[1].view(np.uint32), y[1].view(np.uint32)), f"step {k}: reward bits differ"
        assert np.array_equal(x[2], y[2])
        for u, v in zip(x[3], y[3]):
            assert np.array_equal(u, v), f"step {k}: info differs"
        assert np.array_equal(x[4].view(np.uint64), y[4].view(np.uint64)), f"step {k}: float64 state bits differ"
        assert np.array_equal(x[5], y[5]), f"step {k}: integer state differs"
        finished += int(x[2].sum())
    assert finished > n // 8, "the scenario is supposed to exercise terminations"
    sa, sb = envs[0].read_stats(), envs[1].read_stats()
    for key in sa:
        assert abs(sa[key] - sb[key]) <= 1e-9 * max(1.0, abs(sb[key])), key   # sums of the same terms in another order
    for e in envs:
        e.close()


def test_tracks_on_their_own_streams_are_bitwise_one_stream(torch_cuda, monkeypatch):
    """Batches over several tracks run every track's observation kernels but the first's on a stream of their own
    (fork behind the step kernel, join before the call returns): same outputs as everything on the caller's stream,
    step after step, and the outputs are complete when the caller's stream says so (read right after each call)."""
    torch = torch_cuda
    n = 3000
    envs = []
    for fork in ("1", "0"):
        monkeypatch.setenv("RD_FORK_MAPS", fork)
        envs.append(make_env(torch, tracks=("barcelona", "austria", "columbia"), n_envs=n, action_repeat=4, auto_reset=True,
                             reset_mode="random", seed=12, obs_type="lidar_occupancy"))
    first = [e.reset() for e in envs]
    for key in first[0]:
        assert torch.equal(first[0][key], first[1][key]), f"reset: {key}"
    rng = np.random.RandomState(4)
    for k in range(12):
        a = torch.from_numpy(rng.uniform(-1, 1, (n, 2)).astype(np.float32)).cuda()
        outs = []
        for e in envs:
            obs, rew, done, info = e.step(a)
            outs.append({**{f"obs.{key}": v.clone() for key, v in obs.items()}, "reward": rew.clone(), "done": done.clone()})
        for key in outs[0]:
            assert torch.equal(outs[0][key], outs[1][key]), f"step {k}: {key}"
    for e in envs:
        e.close()
