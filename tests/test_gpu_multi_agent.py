"""GPU parity of the multi-agent worlds against the CPU oracle, through the C ABI: k_step_ma (worlds of
2-4 cars: car-car contact, world-level ActionRepeat / TimeLimit / reset, rank, n_step_progress) and k_lidar's car hits.

Bars: done / flags / lap / rank / opponent masks and the whole integer state bit-exact; LiDAR within 1e-3 m; float64 state
within 1e-5 relative.
"""
import os

import numpy as np
import pytest

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


def make_env(**kw):
    from racing_dreamer_b200 import BatchedRaceEnv, EnvConfig
    return BatchedRaceEnv(EnvConfig(**kw), device="cuda:0")


def make_oracle(env):
    from oracle import Oracle
    return Oracle(env.cfg, env.tracks, env.map_ids, n_threads=THREADS)


def _compare_state(env, orc, k):
    f, i = env.get_state()
    f, i = f.cpu().numpy(), i.cpu().numpy()
    assert np.array_equal(i, orc.i32), f"integer state differs at step {k}"
    scale = np.maximum(np.abs(orc.f64), 1.0)
    assert np.all(np.abs(f - orc.f64) <= DYN_RTOL * scale), f"float64 state differs at step {k}"


@pytest.mark.parametrize("track,A", [("austria", 4), ("treitlstrasse_v2", 2), ("columbia", 3)])
def test_lidar_sees_the_other_cars(torch_cuda, track, A):
    """Teacher-forced poses: worlds of cars a few body lengths apart (and some overlapping / far apart)."""
    torch = torch_cuda
    env = make_env(tracks=(track,), n_envs=A * 8, agents_per_world=A)
    orc = make_oracle(env)
    tm = env.tracks[0]
    rng = np.random.RandomState(21)
    worlds = 512
    anchor = tm.reset_poses[rng.randint(0, len(tm.reset_poses), worlds)]
    poses = np.repeat(anchor, A, axis=0).reshape(worlds, A, 3).copy()
    for a in range(1, A):
        d = rng.uniform(0.0, 4.0, worlds)
        side = rng.uniform(-0.3, 0.3, worlds)
        poses[:, a, 0] += d * np.cos(anchor[:, 2]) - side * np.sin(anchor[:, 2])
        poses[:, a, 1] += d * np.sin(anchor[:, 2]) + side * np.cos(anchor[:, 2])
        poses[:, a, 2] += rng.uniform(-0.8, 0.8, worlds)
    poses[0, 1] = poses[0, 0]                          # two cars on the same spot
    poses[1, 1, :2] = (1e6, -1e6)                      # a car far outside the map
    poses[2, :, 2] = 0.0                               # axis-aligned headings
    poses = poses.reshape(-1, 3)
    got = env.lidar_cast(torch.from_numpy(poses)).cpu().numpy()
    want = orc.lidar_cast(poses)
    assert np.abs(got - want).max() <= LIDAR_TOL_M
    assert np.array_equal(got, want), "car hits are float32 restatements of the same operations: expected identical bits"
    # the cars are really in the picture: the single-agent scan differs
    env1 = make_env(tracks=(track,), n_envs=8)
    alone = env1.lidar_cast(torch.from_numpy(poses)).cpu().numpy()
    assert (got < alone - 0.01).sum() > worlds and np.all(got <= alone)
    env.close(); env1.close()


@pytest.mark.parametrize("track,A,repeat,semantics,tasks", [
    ("austria", 4, 4, "dreamer", None),
    ("treitlstrasse_v2", 2, 8, "dreamer", None),
    ("columbia", 3, 4, "dreamer", ("maximize_progress", "n_step_progress", "max_speed")),
    # the baselines' scenario files: A maximize_progress, B..D n_step_progress, ActionRepeat that never stops early
    ("austria", 4, 4, "baselines", ("maximize_progress", "n_step_progress", "n_step_progress", "n_step_progress")),
])
def test_worlds_closed_loop_vs_oracle(torch_cuda, track, A, repeat, semantics, tasks):
    torch = torch_cuda
    worlds = 384
    n = worlds * A
    env = make_env(tracks=(track,), n_envs=n, agents_per_world=A, agent_tasks=tasks, action_repeat=repeat,
                   repeat_semantics=semantics, auto_reset=True, reset_mode="random_ball", ball_spacing=0.8,
                   time_limit_steps=30, seed=12, laps=1, rescale_actions=(semantics == "dreamer"),
                   clip_actions=(semantics != "dreamer"))
    orc = make_oracle(env)
    o = env.reset()
    r = orc.reset(mode=int(env.cfg.reset_mode))
    _compare_state(env, orc, -1)
    assert np.abs(o["lidar"].cpu().numpy() - r["lidar"]).max() <= LIDAR_TOL_M
    rng = np.random.RandomState(8)
    contacts = dones = 0
    # the car at the back of each world (agent 0) is the fastest, gentle steering: rear-end contacts, not only wall hits
    gain = np.tile(np.array([1.0, 0.15, 0.6, 0.05][:A], np.float32), worlds)
    for k in range(50):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        a[:, 0] = (np.abs(a[:, 0]) * 0.5 + 0.5) * gain
        if semantics == "dreamer":
            a[:, 0] = a[:, 0] * 2 - 1
        a[:, 1] *= 0.35
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"]), k
        assert np.array_equal(info["flags"].cpu().numpy(), ref["flags"]), k
        assert np.array_equal(info["lap"].cpu().numpy(), ref["lap"])
        assert np.array_equal(info["rank"].cpu().numpy(), ref["rank"])
        assert np.array_equal(info["opponent_collisions"].cpu().numpy(), ref["opponents"])
        assert np.allclose(rew.cpu().numpy(), ref["reward"], rtol=DYN_RTOL, atol=1e-5)
        assert np.abs(obs["lidar"].cpu().numpy() - ref["lidar"]).max() <= LIDAR_TOL_M
        _compare_state(env, orc, k)
        contacts += int((ref["opponents"] != 0).sum())
        dones += int(ref["done"].sum())
    assert contacts > 200 and dones > worlds // 4, (contacts, dones)
    # a world is either entirely reset or not at all
    ep = orc.i32[_abi.I_EPISODE].reshape(worlds, A)
    assert np.all(ep == ep[:, :1])
    gs, os_ = env.read_stats(), orc.stats.as_dict()
    for key in gs:
        assert abs(gs[key] - os_[key]) <= 1e-6 * max(1.0, abs(os_[key])), key
    env.close()


def test_worlds_without_auto_reset_freeze_and_masked_reset(torch_cuda):
    torch = torch_cuda
    A, worlds = 2, 64
    n = A * worlds
    env = make_env(tracks=("austria",), n_envs=n, agents_per_world=A, action_repeat=4, auto_reset=False,
                   reset_mode="random_ball", ball_spacing=0.7, time_limit_steps=12, seed=3)
    orc = make_oracle(env)
    env.reset(); orc.reset(mode=int(env.cfg.reset_mode))
    rng = np.random.RandomState(2)
    for k in range(24):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"])
        assert np.array_equal(info["flags"].cpu().numpy(), ref["flags"])
        _compare_state(env, orc, k)
        if k == 8:   # reset by mask: selecting ONE car of a world resets the world
            mask = np.zeros(n, np.uint8)
            mask[1::8] = 1
            o = env.reset(mask=torch.from_numpy(mask).cuda())
            r = orc.reset(mask=mask, mode=int(env.cfg.reset_mode))
            assert np.abs(o["lidar"].cpu().numpy() - r["lidar"]).max() <= LIDAR_TOL_M
            _compare_state(env, orc, k)
    assert ref["done"].all()          # TimeLimit fired for everybody, frozen until reset
    env.close()


def test_single_car_n_step_progress(torch_cuda):
    """agents_per_world = 1 with the n_step_progress task also goes through k_step_ma."""
    torch = torch_cuda
    n = 256
    env = make_env(tracks=("austria",), n_envs=n, task="n_step_progress", n_step_progress=7, action_repeat=4,
                   auto_reset=True, reset_mode="random", time_limit_steps=25, seed=1)
    orc = make_oracle(env)
    env.reset(); orc.reset(mode=int(env.cfg.reset_mode))
    rng = np.random.RandomState(5)
    for k in range(40):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"])
        assert np.allclose(rew.cpu().numpy(), ref["reward"], rtol=DYN_RTOL, atol=1e-5)
        _compare_state(env, orc, k)
    env.close()


def test_host_facing_step_with_worlds(torch_cuda):
    """rd_step_host (numpy in / numpy out, chunked observation kernels) == rd_step for worlds that straddle chunks."""
    torch = torch_cuda
    from racing_dreamer_b200 import EnvConfig
    from racing_dreamer_b200.host import HostSteppedEnv
    A, worlds = 4, 96
    n = A * worlds
    ec = EnvConfig(tracks=("austria",), n_envs=n, agents_per_world=A, action_repeat=4, auto_reset=True,
                   reset_mode="random_ball", ball_spacing=0.8, time_limit_steps=20, seed=6)
    henv = HostSteppedEnv(ec, device="cuda:0", n_shards=5)
    orc = make_oracle(henv.env)
    henv.reset(); orc.reset(mode=int(henv.env.cfg.reset_mode))
    rng = np.random.RandomState(6)
    for k in range(25):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        out = henv.step(a)
        ref = orc.step(a)
        assert np.abs(out["lidar"] - ref["lidar"]).max() <= LIDAR_TOL_M
        for key in ("done", "flags", "lap", "rank", "opponents"):
            assert np.array_equal(out[key], ref[key]), (key, k)
    henv.close()


def test_bad_world_configs_are_rejected(torch_cuda):
    with pytest.raises(RuntimeError, match="agents_per_world"):
        make_env(tracks=("austria",), n_envs=10, agents_per_world=4)
    with pytest.raises(RuntimeError, match="share a map"):
        make_env(tracks=("austria", "barcelona"), n_envs=8, agents_per_world=2, map_ids=[0, 1, 0, 0, 1, 1, 0, 0])
    with pytest.raises(RuntimeError, match="n_step_progress"):
        make_env(tracks=("austria",), n_envs=8, task="n_step_progress", n_step_progress=1000)


def test_world_golden_replay_of_the_reference_wrapper_stack(torch_cuda, golden_dir):
    """One world of four cars through the CUDA path == the UNMODIFIED dict-of-agents wrapper stack of the reference run
    over the one-tick oracle world, step for step (tests/golden/multi_agent_stack_golden.npz)."""
    torch = torch_cuda
    import helpers
    from racing_dreamer_b200 import BatchedRaceEnv
    g = np.load(golden_dir / "multi_agent_stack_golden.npz")
    cfg = helpers.fused_world_config(_abi.default_config(), g)
    env = BatchedRaceEnv(tracks=("austria",), raw_config=cfg, device="cuda:0")

    def step(a):
        obs, rew, done, info = env.step(torch.from_numpy(np.ascontiguousarray(a)).cuda())
        out = {"lidar": obs["lidar"], "pose": obs["pose"], "velocity": obs["velocity"], "speed": obs["speed"],
               "reward": rew, "done": done.to(torch.uint8), "progress": info["progress"], "lap": info["lap"],
               "time": info["time"], "flags": info["flags"], "rank": info["rank"],
               "opponents": info["opponent_collisions"], "occupancy": obs["lidar_occupancy"][..., 0]}
        return {k: v.cpu().numpy() for k, v in out.items()}

    rec = helpers.replay_world(lambda: env.reset(mode="random_ball"), step, g["actions"], g["reset_before"])
    helpers.assert_matches_multi_agent_golden(rec, g, lidar_tol=LIDAR_TOL_M, float_tol=DYN_RTOL)
    env.close()


def test_episode_recorder_for_a_world_of_cars_on_gpu(torch_cuda, golden_dir):
    """§8-f1 x f3 on the CUDA path: the recorder over the host-facing step of a three-car world == what the reference's
    Collect handed to its callbacks (tests/golden/multi_agent_episodes_golden.npz)."""
    from racing_dreamer_b200 import EnvConfig
    from racing_dreamer_b200.episodes import EpisodeRecorder
    from racing_dreamer_b200.host import HostSteppedEnv
    g = np.load(golden_dir / "multi_agent_episodes_golden.npz")
    A = int(g["n_agents"])
    ec = EnvConfig(tracks=("treitlstrasse_v2",), n_envs=A, agents_per_world=A, action_repeat=int(g["action_repeat"]),
                   auto_reset=False, reset_mode="random_ball", time_limit_steps=int(g["duration"]), seed=int(g["seed"]),
                   ball_spacing=float(g["ball_spacing"]))
    env = HostSteppedEnv(ec, device="cuda:0", n_shards=2)
    captured = []
    rec = EpisodeRecorder(env, max_len=int(g["duration"]), callbacks=[captured.append], reset_mode="random_ball")
    rec.reset()
    for t in range(g["actions"].shape[0]):
        rec.step(g["actions"][t])
    assert len(captured) == int(g["n_episodes"])
    for i, eps in enumerate(captured):
        for a, ep in enumerate(eps):
            for k in (str(k) for k in g["keys"]):
                want, got = g[f"ep{i}_{a}_{k}"], ep[k]
                assert got.dtype == want.dtype and got.shape == want.shape, (i, a, k)
                tol = LIDAR_TOL_M if k == "lidar" else DYN_RTOL
                d = np.abs(got.astype(np.float64) - want.astype(np.float64))
                assert np.all(d <= tol * np.maximum(1.0, np.abs(want))), (i, a, k, d.max())
    env.close()
