#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference classes.

Runs only in the build container (needs /root/reference, scipy, Pillow); the fixtures it writes are
committed so that the parity tests can run where /root/reference does not exist (the GPU box).

  occupancy_golden.npz     OccupancyMapObs.step outputs [REF dreamer/wrappers.py:390-408] for seeded poses on
                           three tracks (scipy.ndimage.rotate + PIL resize of THIS image: scipy 1.18 / Pillow 12).
  dreamer_stack_golden.npz BASELINE config 1: the reference wrapper stack of dream.py
                           (RaceCarWrapper -> ActionRepeat(4) -> ReduceActionSpace -> OccupancyMapObs ->
                           FixedResetMode('grid') -> TimeLimit -> Collect) [REF dreamer/dream.py:103-140] over the
                           one-tick oracle env, Columbia, 1000 float32 actions from RandomState(0).
  baselines_stack_golden.npz  the model-free chain's action path: Flatten(clip) -> ActionRepeat(4) with the
                           baselines edge semantics [REF baselines/racing/environment/single_agent.py:31-62].

  gap_follower_golden.npz  the UNMODIFIED follow-the-gap node [REF ros_agent/agents/follow_the_gap/src/agent.py]
                           (rospy stubbed, oracle/ref_ftg.py) fed scan sequences of the oracle env: the drive
                           commands it publishes (steering angle, speed, heading) per scan.

  costmap_golden.npz       the UNMODIFIED map generator [REF docs/maps/costmaps/generate-costmap.py] (skimage/cmapy
                           stubbed over scipy.ndimage, oracle/ref_costmap.py) run on two of the reference's maps:
                           drivable_area, norm_distance_from_start, norm_distance_to_obstacle (+ norm_distance_to).
  raceline_golden.npz      the same generator's race-line layer [REF :280-360,378]: what the UNMODIFIED compute_raceline
                           returned inside run() on f1_aut and columbia_small (sha256 of the float64 layer, track crop,
                           control points).  `python tests/golden/make_golden.py raceline_golden`, ~2 minutes per map.

  multi_agent_stack_golden.npz  one world of four cars (A maximize_progress, B..D n_step_progress as in the baselines'
                           scenario files [REF baselines/scenarios/max_progress/austria.yml]) under the UNMODIFIED
                           dict-of-agents wrapper stack of dream.py (RaceCarWrapper -> ActionRepeat -> ReduceActionSpace
                           -> OccupancyMapObs -> FixedResetMode('random_ball') -> TimeLimit -> Collect)
                           [REF dreamer/dream.py:103-140; dreamer/wrappers.py:107-116,147-154,210-226], reset whenever
                           any agent is done [REF dreamer/tools.py:178-179].

  baselines_chain_golden.npz  the model-free wrap chains end to end: FilterObservation(['lidar']) -> Flatten ->
                           NormalizeObservations (-> InfoToObservation) -> FixedResetMode -> TimeLimit(ticks) ->
                           ActionRepeat [REF baselines/racing/experiments/acme/experiment.py:66-88] with the reference's
                           own classes (gym's FilterObservation / TimeLimit restated in oracle/ref_stubs.py).
  simulate_golden.npz      the reference's driver loop tools.simulate [REF dreamer/tools.py:154-206] (TensorFlow stubbed)
                           over the unmodified dreamer wrapper stack: its per-episode `max_progresses` / `cum_rewards`.

usage: python tests/golden/make_golden.py [name ...]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_stubs  # noqa: E402
from oracle.ref_env import GridMapShim, OracleRaceEnv, make_reference_stack  # noqa: E402
from racing_dreamer_b200 import load_track  # noqa: E402

OUT = Path(__file__).resolve().parent


def occupancy_golden():
    import types
    W = ref_stubs.reference_wrappers()
    rng = np.random.RandomState(1234)
    tracks, poses, images = [], [], []
    for ti, name in enumerate(["austria", "columbia", "treitlstrasse_v2"]):
        tm = load_track(name)

        class Fake:
            agent_ids = ["A"]
            scenario = types.SimpleNamespace(world=types.SimpleNamespace(_maps={"occupancy": GridMapShim(tm)}))
            pose = None

            def step(self, a):
                return {"A": {}}, {"A": 0.0}, {"A": False}, {"A": {"pose": self.pose}}

        env = Fake()
        occ = W.OccupancyMapObs(env)
        p = tm.reset_poses[rng.randint(0, len(tm.reset_poses), 24)].copy()
        p[:, 2] += rng.uniform(-np.pi, np.pi, 24)
        p[:, :2] += rng.uniform(-0.3, 0.3, (24, 2))
        p[0] = tm.start_poses[0]                       # yaw exactly 0
        p[1] = (p[1, 0], p[1, 1], np.pi / 2)           # axis-aligned rotations
        p[2] = (p[2, 0], p[2, 1], -np.pi / 4)
        p[3] = (p[3, 0], p[3, 1], np.pi)
        for q in p:
            env.pose = np.array([q[0], q[1], 0.0, 0.0, 0.0, q[2]])
            o, _, _, _ = occ.step(None)
            img = o["A"]["lidar_occupancy"]
            assert img.shape == (64, 64, 1) and img.dtype == np.uint8 and img.max() <= 1
            images.append(np.packbits(img[..., 0], axis=1))
            poses.append(q)
            tracks.append(ti)
    np.savez_compressed(OUT / "occupancy_golden.npz", track_names=np.array(["austria", "columbia", "treitlstrasse_v2"]),
                        track=np.array(tracks, np.int32), poses=np.array(poses, np.float64),
                        images=np.array(images, np.uint8))
    print("occupancy_golden:", len(poses), "poses")


def dreamer_stack_golden(n_steps=1000, action_repeat=4, duration=125):
    """config 1 [REF dreamer/dream.py:55,57,109: time_limit 2000 ticks / action_repeat]; a shorter TimeLimit (125
    agent steps) so that the time-limit branch also fires inside 1000 steps."""
    tm = load_track("columbia")
    env = make_reference_stack(tm, action_repeat=action_repeat, time_limit_steps=duration, reset_mode="grid")
    actions = np.random.RandomState(0).uniform(-1, 1, (n_steps, 2)).astype(np.float32)
    # damp the steering a little so that episodes last more than a handful of steps
    actions[:, 1] *= 0.35
    rec = {k: [] for k in ("reward", "done", "progress", "lap", "time", "speed", "pose", "velocity", "lidar_sum",
                           "occ_popcount", "reset_before", "wrong_way", "collision")}
    lidar_full, occ_full = [], []
    need_reset = True
    for t in range(n_steps):
        rec["reset_before"].append(need_reset)
        if need_reset:
            obs = env.reset()
            assert obs["A"]["speed"] == 0.0 and not obs["A"]["lidar_occupancy"].any()
        obs, rew, done, info = env.step({"A": actions[t]})
        o, i = obs["A"], info["A"]
        rec["reward"].append(rew["A"])
        rec["done"].append(done["A"])
        rec["progress"].append(i["progress"])
        rec["lap"].append(i["lap"])
        rec["time"].append(i["time"])
        rec["wrong_way"].append(i["wrong_way"])
        rec["collision"].append(i["wall_collision"])
        rec["speed"].append(o["speed"])
        rec["pose"].append(o["pose"])
        rec["velocity"].append(o["velocity"])
        rec["lidar_sum"].append(np.sum(o["lidar"], dtype=np.float64))
        rec["occ_popcount"].append(int(o["lidar_occupancy"].sum()))
        if t % 20 == 0:
            lidar_full.append(o["lidar"])
            occ_full.append(np.packbits(o["lidar_occupancy"][..., 0], axis=1))
        need_reset = bool(done["A"])
    out = {k: np.asarray(v) for k, v in rec.items()}
    assert out["speed"].dtype == np.float32 and out["pose"].dtype == np.float32  # Collect casts [REF wrappers.py:240-250]
    np.savez_compressed(OUT / "dreamer_stack_golden.npz", actions=actions, action_repeat=action_repeat,
                        duration=duration, lidar_every20=np.asarray(lidar_full), occ_every20=np.asarray(occ_full),
                        reward=out["reward"].astype(np.float64), **{k: v for k, v in out.items() if k != "reward"})
    print("dreamer_stack_golden: episodes", int(out["done"].sum()), "collisions", int(out["collision"].sum()),
          "max progress", float((out["lap"] + out["progress"] - 1).max()))


def baselines_stack_golden(n_steps=400, repeat=4):
    B = ref_stubs.reference_baselines_env()
    import gym
    tm = load_track("austria")

    class Single(gym.Wrapper):  # SingleAgentRaceEnv surface over the one-tick oracle env
        def __init__(self, env):
            super().__init__(env)
            self.action_space = env.action_space["A"]
            self.observation_space = env.observation_space["A"]

        def step(self, action):
            o, r, d, i = self.env.step({"A": action})
            return o["A"], r["A"], d["A"], i["A"]

        def reset(self, **kw):
            return self.env.reset(**kw)["A"]

    env = Single(OracleRaceEnv(tm))
    env = B.Flatten(env, flatten_obs=False, flatten_actions=True)
    env = B.ActionRepeat(env, n=repeat)
    actions = np.random.RandomState(5).uniform(-1.4, 1.4, (n_steps, 2)).astype(np.float32)  # exercises the clip
    actions[:, 0] = np.abs(actions[:, 0]) * 0.6 + 0.1
    actions[::7, 0] = 1.3                                   # > 1: clipped by Flatten
    actions[:, 1] *= 0.45
    rec = {k: [] for k in ("reward", "done", "progress", "lap", "time", "reset_before", "collision", "lidar_sum")}
    need_reset = True
    for t in range(n_steps):
        rec["reset_before"].append(need_reset)
        if need_reset:
            env.reset(mode="grid")
        o, r, d, i = env.step(actions[t])
        rec["reward"].append(r); rec["done"].append(d); rec["progress"].append(i["progress"]); rec["lap"].append(i["lap"])
        rec["time"].append(i["time"]); rec["collision"].append(i["wall_collision"])
        rec["lidar_sum"].append(np.sum(o["lidar"].astype(np.float32), dtype=np.float64))
        need_reset = bool(d)
    np.savez_compressed(OUT / "baselines_stack_golden.npz", actions=actions, repeat=repeat,
                        **{k: np.asarray(v) for k, v in rec.items()})
    print("baselines_stack_golden: episodes", int(np.sum(rec["done"])))


class _Single:
    """SingleAgentRaceEnv surface over the one-tick oracle env (racecar_gym's single-agent class is not in tree)."""

    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space["A"]
        self.observation_space = env.observation_space["A"]

    def step(self, action):
        o, r, d, i = self.env.step({"A": action})
        return o["A"], r["A"], d["A"], i["A"]

    def reset(self, **kw):
        return self.env.reset(**kw)["A"]


def baselines_chain_golden(n_steps=300, repeat=4, limit_train=110, limit_test=170):
    """Both model-free wrap chains [REF baselines/racing/experiments/acme/experiment.py:66-88 _wrap_training /
    _wrap_test], every class the reference's own except gym's two (restated, see ref_stubs), over the one-tick oracle
    env on Treitlstrasse (short episodes: collisions AND tick limits fire).  The float32 casts are acme's
    SinglePrecisionWrapper [REF experiment.py:74,86]."""
    B = ref_stubs.reference_baselines_env()
    Cm = ref_stubs.reference_baselines_common()
    tm = load_track("treitlstrasse_v2")
    rng = np.random.RandomState(17)
    actions = rng.uniform(-1.3, 1.3, (n_steps, 2)).astype(np.float32)
    actions[:, 0] = np.abs(actions[:, 0]) * 0.7 + 0.3   # > 1 now and then: clipped by Flatten
    actions[:, 1] *= 0.8
    actions[60:170, 1] *= 0.05         # a stretch of near-straight driving: episodes long enough to hit the tick limit
    out = {"actions": actions, "repeat": repeat, "limit_train": limit_train, "limit_test": limit_test}
    for name, test in (("train", False), ("test", True)):
        env = _Single(OracleRaceEnv(tm))
        env = ref_stubs.FilterObservation(env, filter_keys=["lidar"])
        env = B.Flatten(env, flatten_obs=not test, flatten_actions=True)
        env = B.NormalizeObservations(env)
        if test:
            env = Cm.InfoToObservation(env)
        env = Cm.FixedResetMode(env, mode="grid")
        env = ref_stubs.GymTimeLimit(env, max_episode_steps=limit_test if test else limit_train)
        env = B.ActionRepeat(env, n=repeat)
        rec = {k: [] for k in ("lidar", "reward", "done", "reset_before", "truncated", "progress", "lap", "wrong_way",
                               "wall_collision", "time")}
        need_reset = True
        for t in range(n_steps):
            rec["reset_before"].append(need_reset)
            if need_reset:
                env.reset()
            o, r, d, i = env.step(actions[t])
            lid = o["lidar"] if test else o
            rec["lidar"].append(np.asarray(lid, np.float64).astype(np.float32))
            rec["reward"].append(r); rec["done"].append(d); rec["truncated"].append(bool(i.get("TimeLimit.truncated", False)))
            if test:   # InfoToObservation: every info key shows up as obs['info_<key>'] [REF common.py:31-39]
                assert sorted(k for k in o if k.startswith("info_")) == sorted(f"info_{k}" for k in i if k != "TimeLimit.truncated")
                rec["progress"].append(o["info_progress"]); rec["lap"].append(o["info_lap"]); rec["time"].append(o["info_time"])
                rec["wrong_way"].append(o["info_wrong_way"]); rec["wall_collision"].append(o["info_wall_collision"])
            else:
                rec["progress"].append(i["progress"]); rec["lap"].append(i["lap"]); rec["time"].append(i["time"])
                rec["wrong_way"].append(i["wrong_way"]); rec["wall_collision"].append(i["wall_collision"])
            need_reset = bool(d)
        lid = np.asarray(rec.pop("lidar"))
        out[f"{name}_lidar_every10"] = lid[::10]
        out[f"{name}_lidar_sum"] = lid.sum(1, dtype=np.float64)
        for k, v in rec.items():
            out[f"{name}_{k}"] = np.asarray(v)
        print(f"baselines_chain_golden[{name}]: episodes", int(np.sum(rec["done"])), "truncated", int(np.sum(rec["truncated"])),
              "collisions", int(np.sum(rec["wall_collision"])))
    np.savez_compressed(OUT / "baselines_chain_golden.npz", **out)


def simulate_golden(n_episodes=9, action_repeat=4, duration=40):
    """tools.simulate [REF dreamer/tools.py:154-206], unmodified (TensorFlow stubbed), driving the unmodified dreamer
    wrapper stack over the one-tick oracle env with a scripted agent; what it hands to summarize_collection
    (`max_progresses`, `cum_rewards`) is the per-episode statistic the device-side accumulators must reproduce."""
    import tempfile
    T = ref_stubs.reference_tools()
    tm = load_track("treitlstrasse_v2")
    env = make_reference_stack(tm, action_repeat=action_repeat, time_limit_steps=duration, reset_mode="grid")
    rng = np.random.RandomState(23)
    script = rng.uniform(-1, 1, (4000, 2)).astype(np.float32)
    script[:, 0] = np.abs(script[:, 0]) * 0.6 + 0.2
    script[:, 1] *= 0.3
    used = []

    def agent(obs, done, state):
        a = script[len(used)]
        used.append(a)
        return np.stack([a]), state

    captured = {}
    T.summarize_collection = lambda metrics, *a, **kw: captured.update({k: list(v) for k, v in metrics.items()})
    config = T.AttrDict(action_repeat=action_repeat)
    with tempfile.TemporaryDirectory() as d:
        T.simulate([agent], env, config, Path(d), writer=None, prefix="test", episodes=n_episodes)
    np.savez_compressed(OUT / "simulate_golden.npz", actions=np.asarray(used), action_repeat=action_repeat,
                        duration=duration, n_episodes=n_episodes, max_progresses=np.asarray(captured["progress"]),
                        cum_rewards=np.asarray(captured["return"]))
    print("simulate_golden: steps", len(used), "episodes recorded", len(captured["progress"]), "max_progresses",
          np.round(captured["progress"], 3))


def episodes_golden(n_steps=160, action_repeat=4, duration=45):
    """Episode store wire format: what the reference's Collect wrapper hands to callbacks.save_episodes
    [REF dreamer/wrappers.py:198-250; dreamer/callbacks.py:41-53], captured through a callback on the unmodified
    wrapper stack (Treitlstrasse, reset 'grid', random actions -> collisions and a few time-limit episodes)."""
    from oracle import ref_env
    W = ref_stubs.reference_wrappers()
    tm = load_track("treitlstrasse_v2")
    captured = []
    env = ref_env.OracleRaceEnv(tm)
    env = W.RaceCarWrapper(env, agent_id="A")
    env = W.ActionRepeat(env, action_repeat)
    env = W.ReduceActionSpace(env, low=[0.005, -1.0], high=[1.0, 1.0])
    env = W.OccupancyMapObs(env)
    env = W.FixedResetMode(env, "grid")
    env = W.TimeLimit(env, duration)
    env = W.Collect(env, callbacks=[lambda eps: captured.append({k: np.array(v) for k, v in eps[0].items()})], precision=32)
    actions = np.random.RandomState(9).uniform(-1, 1, (n_steps, 2)).astype(np.float32)
    actions[:, 1] = actions[:, 1] * 0.6 + 0.25      # biased steering: some episodes end in the wall, some at the limit
    reset_before, need_reset = [], True
    for t in range(n_steps):
        reset_before.append(need_reset)
        if need_reset:
            env.reset()
        _, _, done, _ = env.step({"A": actions[t]})
        need_reset = bool(done["A"])
    keys = sorted(captured[0])
    out = {"actions": actions, "reset_before": np.asarray(reset_before), "action_repeat": action_repeat,
           "duration": duration, "n_episodes": len(captured), "keys": np.array(keys)}
    for i, ep in enumerate(captured):
        for k in keys:
            v = ep[k]
            if k == "lidar_occupancy":
                v = np.packbits(v[..., 0], axis=2)
            out[f"ep{i}_{k}"] = v
    np.savez_compressed(OUT / "episodes_golden.npz", **out)
    print("episodes_golden:", len(captured), "episodes, lengths", [len(e["reward"]) for e in captured], "keys", keys,
          "dtypes", {k: str(captured[0][k].dtype) for k in keys})


def gap_follower_golden():
    """Three scan sequences (two closed loops driven by the numpy restatement with steering noise, one of unrelated
    random poses) -> what the reference node publishes after each scan.  Only the forward arc the node reads is stored;
    beams outside it are zero in the reconstructed message."""
    from oracle import Oracle, default_config
    from oracle.gap_follower import GapFollowerOracle, GapFollowerParams
    from oracle.ref_ftg import ReferenceGapFollower
    from racing_dreamer_b200 import _abi
    out = {}
    specs = (("austria", 4, 100, "loop"), ("treitlstrasse_v2", 8, 80, "loop"), ("columbia", 4, 60, "random"))
    for si, (track, R, n_scans, kind) in enumerate(specs):
        tm = load_track(track)
        cfg = default_config()
        cfg.n_envs, cfg.action_repeat, cfg.rescale_actions = 1, R, 0
        orc = Oracle(cfg, [tm], n_threads=1)
        p = GapFollowerParams(dt=R * 0.01)
        s0, s1 = p.arc()
        ref = ReferenceGapFollower(p.angle_min, p.angle_increment, p.n_beams, p.range_max, p.dt)
        mine = GapFollowerOracle(p)
        rng = np.random.RandomState(100 + si)
        scans, cmds = [], []
        if kind == "loop":
            o = orc.reset(mode=_abi.RESET_GRID)
        else:
            poses = tm.reset_poses[rng.randint(0, len(tm.reset_poses), n_scans)].copy()
            poses[:, 2] += rng.uniform(-1.2, 1.2, n_scans)
            poses[:, :2] += rng.uniform(-0.25, 0.25, (n_scans, 2))
            all_scans = orc.lidar_cast(poses)
        veh = cfg.vehicle
        for k in range(n_scans):
            lidar = o["lidar"][0].copy() if kind == "loop" else all_scans[k]
            ros = lidar[::-1].astype(np.float64)
            full = np.zeros_like(ros)
            full[s0:s1 + 1] = ros[s0:s1 + 1]
            pub, sa, sp, hd = ref(full)
            m = mine(full)
            assert m[0] == pub and max(abs(m[1] - sa), abs(m[2] - sp), abs(m[3] - hd)) < 1e-12
            scans.append(lidar[::-1][s0:s1 + 1].astype(np.float32))
            cmds.append((float(pub), sa, sp, hd))
            if kind == "loop":
                v = orc.f64[_abi.S_V, 0]
                vt = sp * 0.5
                motor = min(max(vt * veh.c_drag / veh.a_drive + (vt - v), 0.005), 1.0)
                steer = min(max(sa / (veh.steer_gain * veh.steer_max) + rng.uniform(-0.5, 0.5), -1), 1)
                o = orc.step(commands=np.array([[motor, steer]]))
                if o["done"][0]:
                    break
        out[f"seq{si}_arc_ros"] = np.asarray(scans, np.float32)
        out[f"seq{si}_cmd"] = np.asarray(cmds, np.float64)
        out[f"seq{si}_meta"] = np.array([R, s0, s1, p.n_beams], np.int64)
        print(f"gap_follower_golden: {track} R={R} {len(scans)} scans, published {int(sum(c[0] for c in cmds))}, "
              f"|steer| max {max(abs(c[1]) for c in cmds):.3f}")
    np.savez_compressed(OUT / "gap_follower_golden.npz", n_seq=len(specs), tracks=np.array([s[0] for s in specs]), **out)


def costmap_golden():
    """The generator's three consumed layers for Treitlstrasse_3-U_v2 (where its hard-coded cleared pixel lies ON the
    track) and f1_aut (where it does not).  Stored in the compiled-track layout (crop, integer wavefront distance,
    squared EDT) after checking that this layout reproduces the generator's float arrays bit for bit."""
    from oracle.ref_costmap import reference_layers
    from racing_dreamer_b200 import TRACK_FILES, maps
    out = {"tracks": np.array(["treitlstrasse_v2", "austria"])}
    for name in ("treitlstrasse_v2", "austria"):
        y = ref_stubs.REFERENCE_ROOT / "docs" / "maps" / "maps" / f"{TRACK_FILES[name]}.yaml"
        ref = reference_layers(y)
        tm = maps.compile_track(y, reference_quirks=True)
        assert np.array_equal(ref["drivable_area"], tm.full_drivable())
        assert np.array_equal(ref["norm_distance_from_start"], tm.full_norm_distance_from_start())
        assert np.array_equal(ref["norm_distance_to_obstacle"], tm.full_norm_distance_to_obstacle())
        out[f"{name}_r0c0hw"] = np.array([tm.r0, tm.c0, tm.h, tm.w], np.int64)
        out[f"{name}_drivable"] = np.packbits(tm.drivable, axis=1)
        out[f"{name}_dist"] = tm.dist
        out[f"{name}_dmax"] = tm.dmax
        out[f"{name}_edt_sq"] = tm.edt_sq
        out[f"{name}_start_px"] = np.asarray(ref["grid_starting_position"], np.int64)
        print(f"costmap_golden: {name} crop {tm.h}x{tm.w} dmax {tm.dmax} drivable {int(tm.drivable.sum())} px")
    # the fourth stored layer, norm_distance_to (smoothed distance to the target) [REF generate-costmap.py:227-276,405-420]:
    # the generator's full run() on Treitlstrasse (default settings: use_blurred_factor on).  Stored as the sha256 of the
    # full float64 array (bit-exactness) plus the crop in float32 (diagnostics when the hash differs).
    import hashlib
    y = ref_stubs.REFERENCE_ROOT / "docs" / "maps" / "maps" / f"{TRACK_FILES['treitlstrasse_v2']}.yaml"
    full = reference_layers(y, full_run=True, out_path="/tmp/ref_full_costmap.npz")
    mine = maps.compile_distance_to_target(y, reference_quirks=True)
    assert np.array_equal(full["drivable_area"], mine["drivable_area"])
    assert np.array_equal(full["norm_distance_to"], mine["norm_distance_to"])
    r0, c0, h, w = (int(v) for v in out["treitlstrasse_v2_r0c0hw"])
    ndt = np.ascontiguousarray(full["norm_distance_to"], dtype=np.float64)
    out["treitlstrasse_v2_norm_distance_to_sha256"] = np.array(hashlib.sha256(ndt.tobytes()).hexdigest())
    out["treitlstrasse_v2_norm_distance_to_crop_f32"] = ndt[r0:r0 + h, c0:c0 + w].astype(np.float32)
    print("costmap_golden: norm_distance_to max", float(ndt.max()), "nonzero", int((ndt > 0).sum()))
    np.savez_compressed(OUT / "costmap_golden.npz", **out)


def raceline_golden():
    """the race-line layer [REF docs/maps/costmaps/generate-costmap.py:280-360,378] from the UNMODIFIED
    generator's run() on the two maps with their own settings (f1_aut, columbia_small): sha256 of the full float64 layer,
    the track crop in float32 and the number of non-zero pixels.  ~2 minutes per map."""
    import hashlib
    from oracle.ref_costmap import reference_raceline
    from racing_dreamer_b200 import TRACK_FILES, maps
    out = {}
    for name in ("austria", "columbia"):
        y = ref_stubs.REFERENCE_ROOT / "docs" / "maps" / "maps" / f"{TRACK_FILES[name]}.yaml"
        layer, settings = reference_raceline(y)
        assert settings == maps.generator_settings(TRACK_FILES[name]), (settings, maps.generator_settings(TRACK_FILES[name]))
        mine = maps.compile_raceline(y, reference_quirks=True)
        assert np.array_equal(layer, mine["raceline"]), name
        layer = np.ascontiguousarray(layer, dtype=np.float64)
        rows, cols = np.flatnonzero((layer > 0).any(axis=1)), np.flatnonzero((layer > 0).any(axis=0))
        r0, c0, h, w = int(rows[0]), int(cols[0]), int(rows[-1] - rows[0] + 1), int(cols[-1] - cols[0] + 1)
        out[f"{name}_r0c0hw"] = np.array([r0, c0, h, w], np.int64)
        out[f"{name}_sha256"] = np.array(hashlib.sha256(layer.tobytes()).hexdigest())
        out[f"{name}_crop_f32"] = layer[r0:r0 + h, c0:c0 + w].astype(np.float32)
        out[f"{name}_nonzero"] = np.array(int((layer > 0).sum()))
        out[f"{name}_control_points"] = mine["control_points"]
        print(f"raceline_golden: {name} nonzero {int((layer > 0).sum())} control points {mine['control_points'].shape[0]}")
    np.savez_compressed(OUT / "raceline_golden.npz", **out)


def multi_agent_episodes_golden(n_steps=120, action_repeat=4, duration=15, n_agents=3):
    """SURVEY §8-f1 x f3: what the reference's Collect hands to its callbacks for a world of three cars -- ONE call per
    world episode with a list of per-agent episodes that all end at the same step (discount 0 only for the cars that
    were done themselves) [REF dreamer/wrappers.py:210-238]."""
    from oracle import ref_env
    W = ref_stubs.reference_wrappers()
    tm = load_track("treitlstrasse_v2")
    ids = ["A", "B", "C"][:n_agents]
    captured = []
    env = ref_env.OracleRaceEnv(tm, n_agents=n_agents, seed=23, ball_spacing=0.8)
    env = W.RaceCarWrapper(env, agent_id="A")
    env = W.ActionRepeat(env, action_repeat)
    env = W.ReduceActionSpace(env, low=[0.005, -1.0], high=[1.0, 1.0])
    env = W.FixedResetMode(env, "random_ball")
    env = W.TimeLimit(env, duration)
    env = W.Collect(env, callbacks=[lambda eps: captured.append([{k: np.array(v) for k, v in e.items()} for e in eps])],
                    precision=32)
    rng = np.random.RandomState(23)
    actions = rng.uniform(-1, 1, (n_steps, n_agents, 2)).astype(np.float32)
    actions[:, :, 0] = ((np.abs(actions[:, :, 0]) * 0.5 + 0.5) * np.array([1.0, 0.2, 0.5][:n_agents], np.float32)) * 2 - 1
    actions[:, :, 1] *= 0.3
    reset_before, need_reset = [], True
    for t in range(n_steps):
        reset_before.append(need_reset)
        if need_reset:
            env.reset()
        _, _, done, _ = env.step({i: actions[t, k] for k, i in enumerate(ids)})
        need_reset = any(done.values())
    keys = sorted(captured[0][0])
    out = {"actions": actions, "reset_before": np.asarray(reset_before), "action_repeat": action_repeat, "duration": duration,
           "n_agents": n_agents, "seed": 23, "ball_spacing": 0.8, "n_episodes": len(captured), "keys": np.array(keys)}
    for i, eps in enumerate(captured):
        assert len(eps) == n_agents and len({len(e["reward"]) for e in eps}) == 1
        for a, ep in enumerate(eps):
            for k in keys:
                out[f"ep{i}_{a}_{k}"] = ep[k]
    np.savez_compressed(OUT / "multi_agent_episodes_golden.npz", **out)
    print("multi_agent_episodes_golden:", len(captured), "world episodes, lengths", [len(e[0]["reward"]) for e in captured],
          "terminal discounts", [[float(e["discount"][-1]) for e in eps] for eps in captured][:6])


def multi_agent_stack_golden(n_steps=600, action_repeat=4, duration=18, n_agents=4):
    """the reference wrappers' dict-of-agents semantics (ActionRepeat stops when ANY agent is done and sums
    per agent, TimeLimit sets every done, Collect's casts) over the one-tick multi-car oracle world."""
    tm = load_track("austria")
    tasks = ["maximize_progress"] + ["n_step_progress"] * (n_agents - 1)
    env = make_reference_stack(tm, action_repeat=action_repeat, time_limit_steps=duration, reset_mode="random_ball",
                               n_agents=n_agents, tasks=tasks, seed=17, ball_spacing=0.8)
    ids = ["A", "B", "C", "D"][:n_agents]
    rng = np.random.RandomState(17)
    actions = rng.uniform(-1, 1, (n_steps, n_agents, 2)).astype(np.float32)
    # the car at the back (A) is the fastest, gentle steering: rear-end contacts as well as wall hits and time-outs
    gain = np.array([1.0, 0.15, 0.6, 0.05][:n_agents], np.float32)
    actions[:, :, 0] = ((np.abs(actions[:, :, 0]) * 0.5 + 0.5) * gain) * 2 - 1
    actions[:, :, 1] *= 0.3
    keys = ("reward", "done", "progress", "lap", "time", "speed", "pose", "velocity", "lidar_sum", "occ_popcount",
            "wrong_way", "collision", "opponents", "rank")
    rec = {k: [] for k in keys}
    reset_before, lidar_full, occ_full = [], [], []
    need_reset = True
    for t in range(n_steps):
        reset_before.append(need_reset)
        if need_reset:
            obs = env.reset()
            assert all(obs[i]["speed"] == 0.0 and not obs[i]["lidar_occupancy"].any() for i in ids)
        obs, rew, done, info = env.step({i: actions[t, k] for k, i in enumerate(ids)})
        row = {k: [] for k in keys}
        for k, i in enumerate(ids):
            o, f = obs[i], info[i]
            row["reward"].append(rew[i]); row["done"].append(done[i]); row["progress"].append(f["progress"])
            row["lap"].append(f["lap"]); row["time"].append(f["time"]); row["wrong_way"].append(f["wrong_way"])
            row["collision"].append(f["wall_collision"]); row["rank"].append(f["rank"])
            row["opponents"].append(sum(1 << ids.index(j) for j in f["opponent_collisions"]))
            row["speed"].append(o["speed"]); row["pose"].append(o["pose"]); row["velocity"].append(o["velocity"])
            row["lidar_sum"].append(np.sum(o["lidar"], dtype=np.float64))
            row["occ_popcount"].append(int(o["lidar_occupancy"].sum()))
        for k in keys:
            rec[k].append(row[k])
        if t % 20 == 0:
            lidar_full.append([obs[i]["lidar"] for i in ids])
            occ_full.append([np.packbits(obs[i]["lidar_occupancy"][..., 0], axis=1) for i in ids])
        need_reset = any(done.values())          # [REF dreamer/tools.py:178]
    out = {k: np.asarray(v) for k, v in rec.items()}
    assert out["speed"].dtype == np.float32 and out["pose"].dtype == np.float32
    np.savez_compressed(OUT / "multi_agent_stack_golden.npz", actions=actions, action_repeat=action_repeat,
                        duration=duration, n_agents=n_agents, tasks=np.array(tasks), seed=17, ball_spacing=0.8,
                        reset_before=np.asarray(reset_before), lidar_every20=np.asarray(lidar_full),
                        occ_every20=np.asarray(occ_full), reward=out["reward"].astype(np.float64),
                        **{k: v for k, v in out.items() if k != "reward"})
    print("multi_agent_stack_golden: world episodes", int(out["done"].any(1).sum()), "car contacts",
          int((out["opponents"] != 0).sum()), "wall hits", int(out["collision"].sum()),
          "time-outs", int(out["done"].all(1).sum()))


if __name__ == "__main__":
    assert ref_stubs.available(), "/root/reference is required to regenerate the golden fixtures"
    todo = sys.argv[1:]
    if todo:
        for name in todo:
            globals()[name]()
        sys.exit(0)
    occupancy_golden()
    dreamer_stack_golden()
    baselines_stack_golden()
    episodes_golden()
    multi_agent_stack_golden()
    multi_agent_episodes_golden()
    baselines_chain_golden()
    simulate_golden()
