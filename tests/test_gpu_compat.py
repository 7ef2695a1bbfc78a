"""The reference-facing dict API (racing_dreamer_b200/compat.py) on the GPU, against the golden trajectory recorded from
the UNMODIFIED reference wrapper stack (tests/golden/make_golden.py) -- BASELINE config 1 through the drop-in boundary."""
import numpy as np
import pytest

import helpers
from racing_dreamer_b200 import _abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


def test_reference_env_replays_dreamer_golden(torch_cuda, golden_dir):
    from racing_dreamer_b200.compat import ReferenceEnv
    g = np.load(golden_dir / "dreamer_stack_golden.npz")
    env = ReferenceEnv("columbia", "max_progress", action_repeat=int(g["action_repeat"]),
                       time_limit_steps=int(g["duration"]), reset_mode="grid", device="cuda:0")
    assert env.agent_ids == ["A"] and env.n_agents == 1
    assert env.action_space["A"].shape == (2,)
    assert env.observation_space["A"]["lidar"].shape == (1080,)
    assert env.observation_space["A"]["lidar_occupancy"].shape == (64, 64, 1)
    with pytest.raises(AssertionError, match="Must reset environment"):   # [REF dreamer/wrappers.py:148]
        env.step({"A": np.zeros(2, np.float32)})
    rec = {k: [] for k in ("lidar", "pose", "velocity", "speed", "reward", "done", "progress", "lap", "time", "flags",
                           "occupancy")}
    for t in range(g["actions"].shape[0]):
        if g["reset_before"][t]:
            obs = env.reset()
            assert obs["A"]["speed"] == 0.0 and not obs["A"]["lidar_occupancy"].any()   # [REF wrappers.py:74,410-414]
            assert obs["A"]["lidar_occupancy"].shape == (64, 64, 1) and obs["A"]["lidar"].dtype == np.float32
        obs, rew, done, info = env.step({"A": g["actions"][t].reshape(2)})
        o, i = obs["A"], info["A"]
        rec["lidar"].append(o["lidar"]); rec["pose"].append(o["pose"]); rec["velocity"].append(o["velocity"])
        rec["speed"].append(o["speed"]); rec["occupancy"].append(o["lidar_occupancy"][..., 0])
        rec["reward"].append(np.float32(rew["A"])); rec["done"].append(done["A"])
        rec["progress"].append(np.float32(i["progress"])); rec["lap"].append(i["lap"]); rec["time"].append(np.float32(i["time"]))
        rec["flags"].append((_abi.F_WRONG_WAY if i["wrong_way"] else 0) | (_abi.F_COLLISION if i["wall_collision"] else 0))
    rec = {k: np.stack([np.asarray(x) for x in v]) for k, v in rec.items()}
    helpers.assert_matches_dreamer_golden(rec, g, lidar_tol=1e-3, float_tol=1e-5)
    # the scenario shim OccupancyMapObs reads [REF dreamer/wrappers.py:376,396-399]
    occ = env.scenario.world._maps["occupancy"]
    pr, pc = occ.to_pixel(info["A"]["pose"])
    assert occ._map[pr, pc]                       # the car sits on a drivable pixel
    assert env.scenario.world._config.name
    assert env.render(mode="birds_eye", agent="A").shape == (200, 200, 3)
    assert env.launch_count > 0
    env.close()


def test_tick_env_composes_to_fused_step(torch_cuda):
    """RaceCarGymCompat (one sim tick per step, racecar_gym's interface) driven by a hand-written ActionRepeat loop
    [REF dreamer/wrappers.py:107-116] == one fused ReferenceEnv step."""
    from racing_dreamer_b200.compat import RaceCarGymCompat, ReferenceEnv
    R = 4
    fused = ReferenceEnv("austria", "max_progress", action_repeat=R, time_limit_steps=10 ** 6, reset_mode="grid",
                         occupancy=False, device="cuda:0")
    tick = RaceCarGymCompat("austria", "max_progress", device="cuda:0")
    assert set(tick.action_space["A"].spaces) == {"motor", "steering"}
    fused.reset()
    o = tick.reset(mode="grid")
    assert o["A"]["lidar"].dtype == np.float64 and o["A"]["pose"].shape == (6,)
    rng = np.random.RandomState(5)
    low, high = np.array([0.005, -1.0]), np.array([1.0, 1.0])
    for step in range(150):
        a = rng.uniform(-1, 1, 2).astype(np.float32)
        fo, fr, fd, fi = fused.step({"A": a})
        cmd = (a + 1) / 2 * (high - low) + low          # ReduceActionSpace._normalize [REF dreamer/wrappers.py:129-130]
        total, done = 0.0, False
        for _ in range(R):
            to, tr, td, ti = tick.step({"A": {"motor": cmd[0], "steering": cmd[1]}})
            total += tr["A"]
            done = td["A"]
            if done:
                break
        assert done == fd["A"], step
        assert abs(total - fr["A"]) <= 1e-5 * max(1.0, abs(total))
        assert np.abs(to["A"]["lidar"] - fo["A"]["lidar"]).max() <= 1e-3
        assert ti["A"]["lap"] == fi["A"]["lap"] and ti["A"]["wall_collision"] == fi["A"]["wall_collision"]
        assert np.allclose(ti["A"]["pose"], fi["A"]["pose"], rtol=1e-5, atol=1e-5)
        if done:
            fused.reset()
            tick.reset(mode="grid")
    fused.close()
    tick.close()


@pytest.mark.parametrize("obs_type,n_shards", [("lidar", 4), ("lidar_occupancy", 3)])
def test_host_stepped_env_equals_batched(torch_cuda, obs_type, n_shards):
    """HostSteppedEnv (numpy in / numpy out, sharded streams, slab copies -- the `e2e` call of bench.py) returns exactly
    what one BatchedRaceEnv over the same global env ids returns."""
    torch = torch_cuda
    from racing_dreamer_b200 import BatchedRaceEnv, EnvConfig
    from racing_dreamer_b200.host import HostSteppedEnv
    n = 203                                            # ragged shards
    ec = EnvConfig(tracks=("austria", "treitlstrasse_v2"), n_envs=n, action_repeat=4, obs_type=obs_type, auto_reset=True,
                   reset_mode="random", seed=11, time_limit_steps=7)
    ref = BatchedRaceEnv(ec, device="cuda:0")
    host = HostSteppedEnv(ec, device="cuda:0", n_shards=n_shards)
    ro = ref.reset()
    ho = host.reset()
    assert np.array_equal(ho["lidar"], ro["lidar"].cpu().numpy())
    rng = np.random.RandomState(2)
    for step in range(12):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        obs, rew, done, info = ref.step(torch.from_numpy(a).cuda())
        out = host.step(a)
        assert np.array_equal(out["lidar"], obs["lidar"].cpu().numpy()), step
        assert np.array_equal(out["reward"], rew.cpu().numpy())
        assert np.array_equal(out["done"].astype(bool), done.cpu().numpy())
        assert np.array_equal(out["pose"], obs["pose"].cpu().numpy())
        assert np.array_equal(out["lap"], info["lap"].cpu().numpy())
        assert np.array_equal(out["flags"], info["flags"].cpu().numpy())
        if obs_type == "lidar_occupancy":
            assert np.array_equal(out["occupancy"], obs["lidar_occupancy"].cpu().numpy())
    assert host.d2h_bytes_per_step >= n * 1080 * 4
    s_ref, s_host = ref.read_stats(), host.read_stats()
    assert s_ref["episodes"] == s_host["episodes"] and s_ref["env_steps"] == s_host["env_steps"] == n * 12
    ref.close()
    host.close()


def test_episode_recorder_on_gpu_matches_reference_collect(torch_cuda, golden_dir, tmp_path):
    """episodes recorded from the CUDA env == the ones the reference's Collect wrapper produced."""
    from racing_dreamer_b200 import EnvConfig
    from racing_dreamer_b200.episodes import EpisodeRecorder, count_episodes, save_episodes
    from racing_dreamer_b200.host import HostSteppedEnv
    g = np.load(golden_dir / "episodes_golden.npz")
    ec = EnvConfig(tracks=("treitlstrasse_v2",), n_envs=1, action_repeat=int(g["action_repeat"]), obs_type="lidar_occupancy",
                   auto_reset=False, reset_mode="grid", time_limit_steps=int(g["duration"]))
    env = HostSteppedEnv(ec, device="cuda:0", n_shards=1)
    captured = []
    rec = EpisodeRecorder(env, max_len=int(g["duration"]), reset_mode="grid",
                          callbacks=[lambda eps: captured.append(eps[0]), lambda eps: save_episodes(tmp_path, eps)])
    rec.reset()
    for t in range(g["actions"].shape[0]):
        rec.step(g["actions"][t:t + 1])
    helpers.assert_episodes_match_golden(captured, g, tol=1e-5, lidar_tol=1e-3)
    assert count_episodes(tmp_path)[0] == int(g["n_episodes"])
    env.close()


BASELINES_SCENARIO = """\
world:
  name: austria
agents:
  - id: A
    vehicle: {name: racecar, sensors: [lidar, pose, velocity, acceleration]}
    task:
      task_name: maximize_progress
      params: {laps: 10, time_limit: 180.0, terminate_on_collision: True, collision_reward: -1.0}
  - id: B
    vehicle: {name: racecar, sensors: [lidar, pose, velocity, acceleration], color: red}
    task: {task_name: n_step_progress, params: {n_steps: 10}}
  - id: C
    vehicle: {name: racecar, sensors: [lidar, pose, velocity, acceleration], color: yellow}
    task: {task_name: n_step_progress, params: {n_steps: 10}}
  - id: D
    vehicle: {name: racecar, sensors: [lidar, pose, velocity, acceleration], color: magenta}
    task: {task_name: n_step_progress, params: {n_steps: 10}}
"""


def test_multi_agent_reference_env_replays_the_wrapper_stack_golden(torch_cuda, golden_dir, tmp_path):
    """The four-car scenario of the baselines [REF baselines/scenarios/max_progress/austria.yml] through the dict API
    == the reference's unmodified dict-of-agents wrapper stack (multi_agent_stack_golden.npz)."""
    from racing_dreamer_b200.compat import ReferenceEnv
    g = np.load(golden_dir / "multi_agent_stack_golden.npz")
    yml = tmp_path / "austria.yml"
    yml.write_text(BASELINES_SCENARIO)
    env = ReferenceEnv(scenario=str(yml), action_repeat=int(g["action_repeat"]), time_limit_steps=int(g["duration"]),
                       device="cuda:0", seed=int(g["seed"]), ball_spacing=float(g["ball_spacing"]))
    ids = ["A", "B", "C", "D"]
    assert env.agent_ids == ids and env.n_agents == 4
    assert sorted(env.action_space.spaces) == ids and env.observation_space["C"]["lidar"].shape == (1080,)
    assert env._env.config.reset_mode == "random_ball"          # [REF dreamer/dream.py:105-106]
    keys = ("lidar", "pose", "velocity", "speed", "reward", "done", "progress", "lap", "time", "flags", "occupancy",
            "rank", "opponents")
    rec = {k: [] for k in keys}
    for t in range(g["actions"].shape[0]):
        if g["reset_before"][t]:
            obs = env.reset()
            assert all(obs[i]["speed"] == 0.0 and not obs[i]["lidar_occupancy"].any() for i in ids)
        obs, rew, done, info = env.step({i: g["actions"][t, k] for k, i in enumerate(ids)})
        row = {k: [] for k in keys}
        for k, i in enumerate(ids):
            o, f = obs[i], info[i]
            row["lidar"].append(o["lidar"]); row["pose"].append(o["pose"]); row["velocity"].append(o["velocity"])
            row["speed"].append(o["speed"]); row["occupancy"].append(o["lidar_occupancy"][..., 0])
            row["reward"].append(np.float32(rew[i])); row["done"].append(done[i])
            row["progress"].append(np.float32(f["progress"])); row["lap"].append(f["lap"]); row["time"].append(np.float32(f["time"]))
            row["flags"].append((_abi.F_WRONG_WAY if f["wrong_way"] else 0) | (_abi.F_COLLISION if f["wall_collision"] else 0)
                                | (_abi.F_OPPONENT if f["opponent_collisions"] else 0))
            row["rank"].append(f["rank"])
            row["opponents"].append(sum(1 << ids.index(j) for j in f["opponent_collisions"]))
        for k in keys:
            rec[k].append(np.stack([np.asarray(x) for x in row[k]]))
        if any(done.values()):
            with pytest.raises(AssertionError, match="Must reset environment"):
                env.step({i: np.zeros(2, np.float32) for i in ids})
    rec = {k: np.stack(v) for k, v in rec.items()}
    helpers.assert_matches_multi_agent_golden(rec, g, lidar_tol=1e-3, float_tol=1e-5)
    assert env.render(agent="C").shape == (200, 200, 3)
    env.close()


def test_multi_agent_tick_env_composes_to_fused_step(torch_cuda):
    """RaceCarGymCompat with three cars, driven by the reference's ActionRepeat loop restated by hand
    [REF dreamer/wrappers.py:107-116: stop when ANY agent is done, sum per agent] == the fused multi-car step."""
    from racing_dreamer_b200.compat import RaceCarGymCompat, ReferenceEnv
    R, ids = 4, ["A", "B", "C"]
    kw = dict(n_agents=3, seed=9, ball_spacing=0.8)
    fused = ReferenceEnv("treitlstrasse_v2", "max_progress", action_repeat=R, time_limit_steps=10 ** 6,
                         reset_mode="random_ball", occupancy=False, device="cuda:0", **kw)
    tick = RaceCarGymCompat("treitlstrasse_v2", "max_progress", device="cuda:0", **kw)
    assert tick.agent_ids == ids
    fused.reset()
    tick.reset(mode="random_ball")
    rng = np.random.RandomState(15)
    low, high = np.array([0.005, -1.0]), np.array([1.0, 1.0])
    contacts = 0
    for step in range(200):
        acts = {i: (rng.uniform(-1, 1, 2) * [1.0, 0.3]).astype(np.float32) for i in ids}
        acts["A"][0] = 1.0
        fo, fr, fd, fi = fused.step(acts)
        cmds = {i: (a + 1) / 2 * (high - low) + low for i, a in acts.items()}
        total, dones = {i: 0.0 for i in ids}, {i: False for i in ids}
        for _ in range(R):
            to, tr, dones, ti = tick.step({i: {"motor": c[0], "steering": c[1]} for i, c in cmds.items()})
            total = {i: total[i] + tr[i] for i in ids}
            if any(dones.values()):
                break
        for i in ids:
            assert dones[i] == fd[i], (step, i)
            assert abs(total[i] - fr[i]) <= 1e-5 * max(1.0, abs(total[i]))
            assert np.abs(to[i]["lidar"] - fo[i]["lidar"]).max() <= 1e-3
            assert ti[i]["opponent_collisions"] == fi[i]["opponent_collisions"] and ti[i]["rank"] == fi[i]["rank"]
            contacts += len(fi[i]["opponent_collisions"])
        if any(dones.values()):
            fused.reset()
            tick.reset(mode="random_ball")   # same seed, same episode counters -> the same world
    assert contacts > 0
    fused.close()
    tick.close()


def test_changing_track_env_and_render(torch_cuda):
    """racecar_gym's ChangingTrack envs [REF dreamer/evaluations/make_env.py:6-11; run_evaluation.py:48-49]: every map on
    the device, set_next_env() until the wanted track is current; each track behaves exactly like its own env."""
    from racing_dreamer_b200.compat import ReferenceEnv
    tracks = ["austria", "columbia", "treitlstrasse_v2"]
    multi = ReferenceEnv(tracks, "eval", action_repeat=8, time_limit_steps=50, reset_mode="grid", device="cuda:0")
    rng = np.random.RandomState(1)
    acts = (rng.uniform(-1, 1, (30, 2)) * [1.0, 0.3]).astype(np.float32)
    for want in ["columbia", "treitlstrasse_v2", "austria"]:
        while multi.scenario.world._config.name != multi._tms[tracks.index(want)].name:   # the loop of run_evaluation.py
            multi.set_next_env()
        with pytest.raises(AssertionError, match="Must reset environment"):
            multi.step({"A": acts[0]})
        single = ReferenceEnv(want, "eval", action_repeat=8, time_limit_steps=50, reset_mode="grid", device="cuda:0")
        om, os_ = multi.reset(), single.reset()
        assert np.array_equal(om["A"]["lidar"], os_["A"]["lidar"])
        for a in acts:
            m, s = multi.step({"A": a}), single.step({"A": a})
            assert np.array_equal(m[0]["A"]["lidar"], s[0]["A"]["lidar"])
            assert np.array_equal(m[0]["A"]["lidar_occupancy"], s[0]["A"]["lidar_occupancy"])
            assert m[1] == s[1] and m[2] == s[2] and m[3]["A"]["progress"] == s[3]["A"]["progress"]
            if m[2]["A"]:
                break
        occ = multi.scenario.world._maps["occupancy"]          # OccupancyMapObs would read the CURRENT track's map
        pr, pc = occ.to_pixel(m[3]["A"]["pose"])
        assert occ._map.shape == single.scenario.world._maps["occupancy"]._map.shape
        single.close()
    # order='sequential' [REF baselines/racing/experiments/acme/experiment.py:90-93]: next track at every reset
    seq = ReferenceEnv(tracks[:2], "eval", action_repeat=4, reset_mode="grid", order="sequential", device="cuda:0")
    names = []
    for _ in range(4):
        seq.reset()
        names.append(seq.scenario.world._config.name)
    assert names[0] == names[2] and names[1] == names[3] and names[0] != names[1]
    # render: both views, the car in the middle (red), the track white under it
    for mode in ("birds_eye", "follow"):
        img = seq.render(mode=mode, agent="A")
        assert img.shape == (200, 200, 3) and img.dtype == np.uint8
        assert tuple(img[100, 100]) == (255, 0, 0) and (img == 255).all(axis=2).sum() > 500
    with pytest.raises(ValueError):
        seq.render(mode="nope")
    seq.close()
    multi.close()


def test_single_agent_env_shape_of_the_baselines(torch_cuda, golden_dir):
    """SingleAgentRaceCompat (racecar_gym's SingleAgentRaceEnv surface) under the baselines' action chain restated by hand
    -- Flatten's clip to [-1, 1] and ActionRepeat(4) that does not test the first tick's done
    [REF baselines/racing/environment/single_agent.py:31-62] -- replays the fixture recorded from the unmodified classes;
    with a list of tracks every reset moves on to the next one [REF baselines/racing/experiments/acme/experiment.py:90-93]."""
    from racing_dreamer_b200 import SingleAgentRaceCompat
    g = np.load(golden_dir / "baselines_stack_golden.npz")
    R = int(g["repeat"])
    env = SingleAgentRaceCompat("austria", "max_progress", device="cuda:0")
    assert set(env.action_space.spaces) == {"motor", "steering"} and env.observation_space["lidar"].shape == (1080,)
    rec = {k: [] for k in ("reward", "done", "progress", "lap", "time", "flags", "lidar")}
    for t in range(g["actions"].shape[0]):
        if g["reset_before"][t]:
            env.reset(mode="grid")
        a = np.clip(g["actions"][t], -1.0, 1.0)                                   # Flatten.step
        act = {"motor": a[0], "steering": a[1]}
        obs, total, done, info = env.step(act)                                   # ActionRepeat.step
        for _ in range(R - 1):
            obs, r, done, info = env.step(act)
            total += r
            if done:
                break
        rec["reward"].append(np.float32(total)); rec["done"].append(done); rec["progress"].append(np.float32(info["progress"]))
        rec["lap"].append(info["lap"]); rec["time"].append(np.float32(info["time"]))
        rec["flags"].append(_abi.F_COLLISION if info["wall_collision"] else 0)
        rec["lidar"].append(obs["lidar"].astype(np.float32))
    rec = {k: np.stack([np.asarray(x) for x in v]) for k, v in rec.items()}
    helpers.assert_matches_baselines_golden(rec, g, float_tol=1e-5, lidar_tol=1e-3)
    env.close()
    seq = SingleAgentRaceCompat(["austria", "columbia"], "max_progress", device="cuda:0", order="sequential")
    names = []
    for _ in range(4):
        seq.reset(mode="grid")
        names.append(seq.scenario.world._config.name)
    assert names[0] == names[2] != names[1] == names[3]
    seq.close()
