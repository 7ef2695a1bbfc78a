"""Shared replay logic: drive a fused env (CPU oracle or the CUDA path) through a golden trajectory that was
recorded from the reference wrapper stack, and collect the same per-step quantities."""
import numpy as np

from racing_dreamer_b200 import _abi


def fused_dreamer_config(cfg, action_repeat, duration, occupancy=True):
    """rd_config equivalent of dream.py's env stack [REF dreamer/dream.py:103-140] with FixedResetMode('grid')."""
    cfg.n_envs = 1
    cfg.action_repeat = int(action_repeat)
    cfg.repeat_semantics = _abi.REPEAT_DREAMER
    cfg.rescale_actions = 1
    cfg.clip_actions = 0
    cfg.time_limit_steps = int(duration)
    cfg.auto_reset = 0
    cfg.reset_mode = _abi.RESET_GRID
    cfg.obs_flags = _abi.OBS_LIDAR | (_abi.OBS_OCCUPANCY if occupancy else 0)
    return cfg


def fused_baselines_config(cfg, repeat):
    """Flatten(clip) -> ActionRepeat(n) of the model-free chain [REF baselines/racing/environment/single_agent.py:31-62]."""
    cfg.n_envs = 1
    cfg.action_repeat = int(repeat)
    cfg.repeat_semantics = _abi.REPEAT_BASELINES
    cfg.rescale_actions = 0
    cfg.clip_actions = 1
    cfg.time_limit_steps = 0
    cfg.auto_reset = 0
    cfg.obs_flags = _abi.OBS_LIDAR
    return cfg


def replay(reset_fn, step_fn, actions, reset_before):
    """reset_fn() -> None ; step_fn(a[1,2] f32) -> dict of numpy arrays (batch 1).  Returns dict of stacked records."""
    rec = {}
    for t in range(actions.shape[0]):
        if reset_before[t]:
            reset_fn()
        out = step_fn(actions[t:t + 1])
        for k, v in out.items():
            if v is None:
                continue
            rec.setdefault(k, []).append(np.array(v[0]))
    return {k: np.stack(v) for k, v in rec.items()}


def assert_matches_dreamer_golden(rec, g, lidar_tol=0.0, float_tol=0.0):
    """Flags/occupancy bit-exact always; float quantities within the given tolerance (0 = bit-exact)."""
    def close(a, b, tol):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if tol == 0.0:
            assert np.array_equal(a, b), f"max |diff| {np.abs(a - b).max()}"
        else:
            assert np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))), f"max |diff| {np.abs(a - b).max()}"
    assert np.array_equal(rec["done"].astype(bool), g["done"])
    assert np.array_equal(rec["lap"], g["lap"])
    fl = rec["flags"]
    assert np.array_equal((fl & _abi.F_COLLISION) != 0, g["collision"])
    assert np.array_equal((fl & _abi.F_WRONG_WAY) != 0, g["wrong_way"])
    occ = rec["occupancy"].reshape(len(fl), 64, 64)
    assert np.array_equal(occ.reshape(len(occ), -1).sum(1), g["occ_popcount"])
    assert np.array_equal(np.packbits(occ[::20], axis=2), g["occ_every20"])
    close(rec["reward"], g["reward"].astype(np.float32), float_tol)
    close(rec["progress"], g["progress"].astype(np.float32), float_tol)
    close(rec["time"], g["time"].astype(np.float32), float_tol)
    close(rec["speed"], g["speed"], float_tol)
    close(rec["pose"], g["pose"], float_tol)
    close(rec["velocity"], g["velocity"], float_tol)
    if lidar_tol == 0.0:
        assert np.array_equal(rec["lidar"][::20], g["lidar_every20"])
    else:
        assert np.abs(rec["lidar"][::20] - g["lidar_every20"]).max() <= lidar_tol
    assert np.abs(rec["lidar"].sum(1, dtype=np.float64) - g["lidar_sum"]).max() <= max(lidar_tol * 1080, 0.0)


def assert_matches_baselines_golden(rec, g, float_tol=0.0, lidar_tol=0.0):
    assert np.array_equal(rec["done"].astype(bool), g["done"])
    assert np.array_equal(rec["lap"], g["lap"])
    assert np.array_equal((rec["flags"] & _abi.F_COLLISION) != 0, g["collision"])
    for k in ("reward", "progress", "time"):
        a, b = rec[k].astype(np.float64), g[k].astype(np.float32).astype(np.float64)
        if float_tol == 0.0:
            assert np.array_equal(a, b), k
        else:
            assert np.all(np.abs(a - b) <= float_tol * np.maximum(1.0, np.abs(b))), k
    assert np.abs(rec["lidar"].sum(1, dtype=np.float64) - g["lidar_sum"]).max() <= lidar_tol * 1080


class OracleHostEnv:
    """The CPU oracle behind the host-facing env interface (reset(mask, mode) / step(actions) -> dict of numpy arrays),
    so that host-side logic written for HostSteppedEnv (e.g. the episode recorder) can be tested without a GPU."""

    def __init__(self, cfg, tracks, map_ids=None, n_threads=1):
        from oracle import Oracle
        self.orc = Oracle(cfg, tracks, map_ids, n_threads=n_threads)
        self.cfg = self.orc.cfg
        self.n = self.orc.n

    def _out(self, out):
        d = {k: v for k, v in out.items() if v is not None and k != "reward64"}
        if "occupancy" in d:
            d["occupancy"] = d["occupancy"][..., None]
        return d

    def reset(self, mask=None, mode=None):
        m = _abi.RESET_MODES[mode] if mode is not None else int(self.cfg.reset_mode)
        return self._out(self.orc.reset(mask=mask, mode=m))

    def step(self, actions):
        return self._out(self.orc.step(actions))


def assert_episodes_match_golden(episodes, g, tol=1e-6, lidar_tol=1e-3):
    """episodes: list of dicts as handed to Collect's callbacks; g: tests/golden/episodes_golden.npz."""
    assert len(episodes) == int(g["n_episodes"])
    keys = [str(k) for k in g["keys"]]
    for i, ep in enumerate(episodes):
        assert sorted(ep) == keys, (sorted(ep), keys)
        for k in keys:
            want = g[f"ep{i}_{k}"]
            got = ep[k]
            if k == "lidar_occupancy":
                assert got.dtype == np.uint8 and got.shape[1:] == (64, 64, 1)
                assert np.array_equal(np.packbits(got[..., 0], axis=2), want), f"episode {i} {k}"
                continue
            assert got.dtype == want.dtype and got.shape == want.shape, f"episode {i} {k}: {got.dtype}{got.shape} vs {want.dtype}{want.shape}"
            t = lidar_tol if k == "lidar" else tol
            d = np.abs(got.astype(np.float64) - want.astype(np.float64))
            assert np.all(d <= t * np.maximum(1.0, np.abs(want))), f"episode {i} {k}: max diff {d.max()}"


# ---------------------------------------------------------------------------------------------- multi-agent worlds
def fused_world_config(cfg, g, occupancy=True):
    """rd_config equivalent of the reference's dict-of-agents stack recorded in multi_agent_stack_golden.npz: ONE world of
    n_agents cars, manual reset whenever any car is done [REF dreamer/tools.py:178-179]."""
    A = int(g["n_agents"])
    cfg.n_envs = A
    cfg.agents_per_world = A
    for a, name in enumerate(g["tasks"]):
        cfg.agent_task[a] = _abi.TASKS[str(name)]
    cfg.action_repeat = int(g["action_repeat"])
    cfg.repeat_semantics = _abi.REPEAT_DREAMER
    cfg.rescale_actions = 1
    cfg.clip_actions = 0
    cfg.time_limit_steps = int(g["duration"])
    cfg.auto_reset = 0
    cfg.reset_mode = _abi.RESET_RANDOM_BALL
    cfg.seed = int(g["seed"])
    cfg.ball_spacing = float(g["ball_spacing"])
    cfg.obs_flags = _abi.OBS_LIDAR | (_abi.OBS_OCCUPANCY if occupancy else 0)
    return cfg


def replay_world(reset_fn, step_fn, actions, reset_before):
    """step_fn(a[A,2] f32) -> dict of numpy arrays with leading dim A.  Returns dict of [T, A, ...] records."""
    rec = {}
    for t in range(actions.shape[0]):
        if reset_before[t]:
            reset_fn()
        out = step_fn(actions[t])
        for k, v in out.items():
            if v is not None:
                rec.setdefault(k, []).append(np.array(v))
    return {k: np.stack(v) for k, v in rec.items()}


def assert_matches_multi_agent_golden(rec, g, lidar_tol=0.0, float_tol=0.0):
    def close(a, b, tol, what):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if tol == 0.0:
            assert np.array_equal(a, b), f"{what}: max |diff| {np.abs(a - b).max()}"
        else:
            assert np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))), f"{what}: max |diff| {np.abs(a - b).max()}"
    T, A = g["done"].shape
    assert np.array_equal(rec["done"].astype(bool), g["done"])
    assert np.array_equal(rec["lap"], g["lap"])
    assert np.array_equal(rec["rank"], g["rank"])
    assert np.array_equal(rec["opponents"], g["opponents"])
    fl = rec["flags"]
    assert np.array_equal((fl & _abi.F_COLLISION) != 0, g["collision"])
    assert np.array_equal((fl & _abi.F_WRONG_WAY) != 0, g["wrong_way"])
    assert np.array_equal((fl & _abi.F_OPPONENT) != 0, g["opponents"] != 0)
    occ = rec["occupancy"].reshape(T, A, 64, 64)
    assert np.array_equal(occ.reshape(T, A, -1).sum(2), g["occ_popcount"])
    assert np.array_equal(np.packbits(occ[::20], axis=3), g["occ_every20"])
    close(rec["reward"], g["reward"].astype(np.float32), float_tol, "reward")
    close(rec["progress"], g["progress"].astype(np.float32), float_tol, "progress")
    close(rec["time"], g["time"].astype(np.float32), float_tol, "time")
    for k in ("speed", "pose", "velocity"):
        close(rec[k], g[k], float_tol, k)
    if lidar_tol == 0.0:
        assert np.array_equal(rec["lidar"][::20], g["lidar_every20"])
    else:
        assert np.abs(rec["lidar"][::20] - g["lidar_every20"]).max() <= lidar_tol
    assert np.abs(rec["lidar"].sum(2, dtype=np.float64) - g["lidar_sum"]).max() <= max(lidar_tol * 1080, 0.0)


# ---------------------------------------------------------------------------------------------- model-free chain (a11)
def fused_baselines_chain_config(cfg, g, test: bool):
    """rd_config equivalent of the model-free wrap chains [REF baselines/racing/experiments/acme/experiment.py:66-88]:
    Flatten(clip) -> NormalizeObservations -> FixedResetMode('grid') -> TimeLimit(ticks) -> ActionRepeat."""
    cfg.n_envs = 1
    cfg.action_repeat = int(g["repeat"])
    cfg.repeat_semantics = _abi.REPEAT_BASELINES
    cfg.rescale_actions = 0
    cfg.clip_actions = 1
    cfg.time_limit_steps = 0
    cfg.time_limit_ticks = int(g["limit_test"] if test else g["limit_train"])
    cfg.auto_reset = 0
    cfg.reset_mode = _abi.RESET_GRID
    cfg.obs_flags = _abi.OBS_LIDAR | _abi.OBS_NORM_BASELINES
    return cfg


def assert_matches_baselines_chain_golden(rec, g, name, lidar_tol=0.0, float_tol=0.0):
    """rec: replay() records of the fused env; g: baselines_chain_golden.npz; name: 'train' | 'test'."""
    assert np.array_equal(rec["done"].astype(bool), g[f"{name}_done"])
    assert np.array_equal(rec["lap"], g[f"{name}_lap"])
    fl = rec["flags"]
    assert np.array_equal((fl & _abi.F_COLLISION) != 0, g[f"{name}_wall_collision"])
    assert np.array_equal((fl & _abi.F_WRONG_WAY) != 0, g[f"{name}_wrong_way"])
    for k in ("reward", "progress", "time"):
        a, b = rec[k].astype(np.float64), g[f"{name}_{k}"].astype(np.float32).astype(np.float64)
        if float_tol == 0.0:
            assert np.array_equal(a, b), k
        else:
            assert np.all(np.abs(a - b) <= float_tol * np.maximum(1.0, np.abs(b))), k
    lid = rec["lidar"]
    assert lid.dtype == np.float32 and lid.min() >= 0.0 and lid.max() <= 1.0     # NormalizeObservations: [0, 1]
    if lidar_tol == 0.0:
        assert np.array_equal(lid[::10], g[f"{name}_lidar_every10"])
    else:
        assert np.abs(lid[::10] - g[f"{name}_lidar_every10"]).max() <= lidar_tol
    assert np.abs(lid.sum(1, dtype=np.float64) - g[f"{name}_lidar_sum"]).max() <= max(lidar_tol * 1080, 0.0)


# ---------------------------------------------------------------------------------------------- tools.simulate (a12)
def simulate_statistics(reset_fn, step_fn, read_stats_fn, g):
    """Drives a fused single env (TimeLimit, ActionRepeat inside) through simulate_golden.npz's action script with
    simulate's reset rule (reset when done) and turns the device-side per-episode statistics into the two lists
    tools.simulate hands to summarize_collection [REF dreamer/tools.py:154-206]:
      cum_rewards[i]    = return of episode i                                   -> rd_stats.return_sum per episode
      max_progresses[i] = max(episode_progresses) at the reset after episode i; the reference never clears that list, so
                          it is the running maximum over every step so far      -> cummax of rd_stats.max_progress_sum
    Both lists are appended at the NEXT reset, so the last episode of a call is not in them."""
    returns, maxima = [], []
    need_reset = True
    for a in g["actions"]:
        if need_reset:
            reset_fn()
        out = step_fn(a[None])
        need_reset = bool(out["done"][0])
        if need_reset:
            st = read_stats_fn()
            assert st["episodes"] == 1.0
            returns.append(st["return_sum"])
            maxima.append(st["max_progress_sum"])
    assert need_reset and len(returns) == int(g["n_episodes"])
    return np.asarray(returns[:-1]), np.maximum.accumulate(np.asarray(maxima))[:-1]
