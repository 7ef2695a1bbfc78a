"""Episode store: EpisodeRecorder == the reference's Collect wrapper, file format == save_episodes,
sampler == load_episodes.  CPU: the recorder runs over the oracle behind the host-env interface."""
import numpy as np

import helpers
from racing_dreamer_b200 import _abi, load_track
from racing_dreamer_b200.episodes import EpisodeRecorder, count_episodes, load_episodes, save_episodes


def _cfg(n, action_repeat, duration, occupancy=True):
    from oracle import default_config
    cfg = helpers.fused_dreamer_config(default_config(), action_repeat, duration, occupancy)
    cfg.n_envs = n
    return cfg


def test_recorder_reproduces_reference_collect(golden_dir):
    g = np.load(golden_dir / "episodes_golden.npz")
    env = helpers.OracleHostEnv(_cfg(1, int(g["action_repeat"]), int(g["duration"])), [load_track("treitlstrasse_v2")])
    captured = []
    rec = EpisodeRecorder(env, max_len=int(g["duration"]), callbacks=[lambda eps: captured.append(eps[0])], reset_mode="grid")
    for t in range(g["actions"].shape[0]):
        if g["reset_before"][t]:
            if t == 0:
                rec.reset()
        rec.step(g["actions"][t:t + 1])            # finished envs are flushed and reset inside
    helpers.assert_episodes_match_golden(captured, g)
    first = captured[0]
    assert first["reward"][0] == 0 and first["discount"][0] == 1 and first["progress"][0] == -1   # [REF wrappers.py:228-238]
    assert not first["lidar_occupancy"][0].any() and first["speed"][0] == 0
    assert first["discount"][-1] == 0                # terminal transition


def test_batched_recorder_files_and_sampler(tmp_path):
    n, R, T = 6, 4, 12
    env = helpers.OracleHostEnv(_cfg(n, R, T, occupancy=False), [load_track("austria")])
    written = []
    rec = EpisodeRecorder(env, max_len=T, callbacks=[lambda eps: written.extend(save_episodes(tmp_path, eps))],
                          reset_mode="random")
    rec.reset()
    rng = np.random.RandomState(1)
    steps = 40
    for _ in range(steps):
        out = rec.step(rng.uniform(-1, 1, (n, 2)).astype(np.float32))
        assert out["lidar"].shape == (n, 1080)
    assert rec.episodes_done == len(written) >= n * (steps // T)
    n_eps, n_steps = count_episodes(tmp_path)
    assert n_eps == len(written)
    total_rows = 0
    for path in written:
        ep = np.load(path)
        length = int(path.stem.rsplit("-", 1)[-1])                     # {timestamp}-{uuid}-{length}.npz
        assert set(ep.keys()) == {"lidar", "pose", "velocity", "speed", "action", "reward", "discount", "progress", "time"}
        assert all(len(ep[k]) == length for k in ep.keys()) and 2 <= length <= T + 1
        assert ep["discount"][-1] == 0 and np.all(ep["discount"][:-1] == 1)
        assert ep["lidar"].dtype == np.float32 and ep["action"].dtype == np.float32
        assert np.all(np.diff(ep["time"][1:]) > 0)                      # time restarts at every episode
        total_rows += length
    assert n_steps == total_rows - n_eps
    chunks = load_episodes(tmp_path, rescan=8, length=3, seed=0)
    for _ in range(20):
        c = next(chunks)
        assert all(len(v) == 3 for v in c.values())
    a = [next(load_episodes(tmp_path, 4, length=3, seed=5))["reward"] for _ in range(2)]
    assert np.array_equal(a[0], a[1])                                   # same seed -> same stream


def test_recorder_reproduces_reference_collect_for_a_world_of_cars(golden_dir):
    """§8-f1 x f3: one world of three cars -- every car's episode ends when any car is done, one callback per world with
    the list of per-agent episodes, discount 0 only for the cars that were done themselves
    [REF dreamer/wrappers.py:210-238]; fixture recorded from the unmodified Collect (multi_agent_episodes_golden)."""
    from oracle import default_config
    g = np.load(golden_dir / "multi_agent_episodes_golden.npz")
    A = int(g["n_agents"])
    cfg = helpers.fused_dreamer_config(default_config(), int(g["action_repeat"]), int(g["duration"]), occupancy=False)
    cfg.n_envs = A
    cfg.agents_per_world = A
    cfg.reset_mode = _abi.RESET_RANDOM_BALL
    cfg.seed = int(g["seed"])
    cfg.ball_spacing = float(g["ball_spacing"])
    env = helpers.OracleHostEnv(cfg, [load_track("treitlstrasse_v2")])
    captured = []
    rec = EpisodeRecorder(env, max_len=int(g["duration"]), callbacks=[captured.append], reset_mode="random_ball")
    assert rec.agents == A
    rec.reset()
    for t in range(g["actions"].shape[0]):
        rec.step(g["actions"][t])
    assert len(captured) == int(g["n_episodes"])
    keys = [str(k) for k in g["keys"]]
    for i, eps in enumerate(captured):
        assert len(eps) == A
        for a, ep in enumerate(eps):
            assert sorted(ep) == keys
            for k in keys:
                want, got = g[f"ep{i}_{a}_{k}"], ep[k]
                assert got.dtype == want.dtype and got.shape == want.shape, (i, a, k)
                assert np.array_equal(got, want), (i, a, k)
    assert any(eps[0]["discount"][-1] == 0 and eps[-1]["discount"][-1] == 1 for eps in captured)   # mixed endings occur
