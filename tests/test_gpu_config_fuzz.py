"""Seeded sweep over the configuration space of the fused step on the GPU against the oracle: tracks, batch shapes, beam
counts, repeat counts and semantics, tasks, reset modes, time limits, noise, normalisation, worlds.  Each case is a short
closed loop with random actions; flags / laps / dones and the integer state must be bit-exact, scans within 1e-3 m, float
state within 1e-5 relative (the bars of BASELINE.json's north_star)."""
import os

import numpy as np
import pytest

from racing_dreamer_b200 import _abi

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1
TRACKS = ["austria", "columbia", "treitlstrasse_v2", "barcelona", "gbr", "circle_cw"]


def _case(seed):
    r = np.random.RandomState(1000 + seed)
    A = int(r.choice([1, 1, 1, 2, 3, 4]))
    tracks = tuple(r.choice(TRACKS, size=int(r.choice([1, 1, 2, 3])), replace=False))
    worlds = int(r.choice([1, 7, 33, 96]))
    sem = str(r.choice(["dreamer", "dreamer", "baselines"]))
    tasks_all = ["maximize_progress", "max_speed", "n_step_progress"]
    kw = dict(
        tracks=tracks, n_envs=worlds * A, agents_per_world=A,
        n_beams=int(r.choice([1080, 1080, 360, 97])), action_repeat=int(r.choice([1, 2, 4, 8])), repeat_semantics=sem,
        rescale_actions=bool(sem == "dreamer"), clip_actions=bool(sem != "dreamer" or r.rand() < 0.3),
        task=str(r.choice(tasks_all)) if A == 1 else "maximize_progress",
        agent_tasks=tuple(r.choice(tasks_all, size=A)) if A > 1 else None,
        n_step_progress=int(r.choice([1, 3, 10, 32])), laps=int(r.choice([1, 2, 10])),
        terminate_on_collision=bool(r.rand() < 0.8), progress_abs=bool(r.rand() < 0.2),
        collision_reward=float(r.choice([-1.0, -5.0, 0.0])), frame_reward=float(r.choice([0.0, -0.01])),
        n_checkpoints=int(r.choice([1, 4, 20, 50])), time_limit_steps=int(r.choice([0, 5, 17])),
        time_limit=float(r.choice([180.0, 0.35])), auto_reset=bool(r.rand() < 0.7),
        reset_mode=str(r.choice(["grid", "random", "random_bidirectional", "random_ball"])),
        lidar_noise=float(r.choice([0.0, 0.0, 0.03])), normalize_lidar=bool(r.rand() < 0.3),
        time_limit_ticks=int(r.choice([0, 0, 13, 40])),
        obs_type=str(r.choice(["lidar", "lidar", "lidar_occupancy"])) if worlds * A <= 132 else "lidar",
        ball_spacing=float(r.choice([0.7, 1.5])), seed=int(r.randint(0, 2 ** 31)), env_id_offset=int(r.choice([0, 4096])) * A)
    if not kw["normalize_lidar"] and r.rand() < 0.3:   # the model-free chain's NormalizeObservations
        kw["normalize_obs"] = "baselines"
    return kw


@pytest.mark.parametrize("seed", range(int(os.environ.get("RD_FUZZ_SEEDS", "32"))))
def test_random_configuration_vs_oracle(seed):
    import torch
    from oracle import Oracle
    from racing_dreamer_b200 import BatchedRaceEnv, EnvConfig
    torch.cuda.set_device(0)
    kw = _case(seed)
    env = BatchedRaceEnv(EnvConfig(**kw), device="cuda:0")
    orc = Oracle(env.cfg, env.tracks, env.map_ids, n_threads=THREADS)
    n = env.n
    o, r = env.reset(), orc.reset(mode=int(env.cfg.reset_mode))
    tol = 1e-3 if not (kw["normalize_lidar"] or kw.get("normalize_obs")) else 1e-3 / 15.0
    assert np.abs(o["lidar"].cpu().numpy() - r["lidar"]).max() <= tol, kw
    rng = np.random.RandomState(seed)
    for k in range(30):
        a = rng.uniform(-1.3, 1.3, (n, 2)).astype(np.float32)
        if not kw["auto_reset"] and k % 9 == 8:          # frozen envs: reset the done ones (whole worlds) by mask
            mask = (orc.i32[_abi.I_FLAGS] & _abi.F_NEEDS_RESET) != 0
            if mask.any():
                env.reset(mask=torch.from_numpy(mask.astype(np.uint8)).cuda())
                orc.reset(mask=mask.astype(np.uint8), mode=int(env.cfg.reset_mode))
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref = orc.step(a)
        assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"]), (k, kw)
        assert np.array_equal(info["flags"].cpu().numpy(), ref["flags"]), (k, kw)
        assert np.array_equal(info["lap"].cpu().numpy(), ref["lap"]), (k, kw)
        assert done.dtype == torch.bool and info["wrong_way"].dtype == torch.bool
        assert np.array_equal(info["wrong_way"].cpu().numpy(), (ref["flags"] & _abi.F_WRONG_WAY) != 0), (k, kw)
        assert np.array_equal(info["wall_collision"].cpu().numpy(), (ref["flags"] & _abi.F_COLLISION) != 0), (k, kw)
        for key in ("pose", "velocity"):
            assert np.allclose(obs[key].cpu().numpy(), ref[key], rtol=1e-5, atol=1e-6), (k, key, kw)
        assert np.array_equal(info["rank"].cpu().numpy(), ref["rank"]) or kw["agents_per_world"] == 1, (k, kw)
        assert np.allclose(rew.cpu().numpy(), ref["reward"], rtol=1e-5, atol=2e-5), (k, kw)
        assert np.abs(obs["lidar"].cpu().numpy() - ref["lidar"]).max() <= tol, (k, kw)
        if "lidar_occupancy" in obs:
            assert np.array_equal(obs["lidar_occupancy"].cpu().numpy()[..., 0], ref["occupancy"]), (k, kw)
        f, i = env.get_state()
        assert np.array_equal(i.cpu().numpy(), orc.i32), (k, kw)
        assert np.all(np.abs(f.cpu().numpy() - orc.f64) <= 1e-5 * np.maximum(np.abs(orc.f64), 1.0)), (k, kw)
    gs, os_ = env.read_stats(), orc.read_stats()
    for key in ("episodes", "collisions", "laps_completed", "env_steps", "timeouts", "length_sum"):
        assert gs[key] == os_[key], (key, gs, os_, kw)
    for key in ("return_sum", "progress_sum", "max_progress_sum"):
        assert abs(gs[key] - os_[key]) <= 1e-6 * max(1.0, abs(os_[key])), (key, gs, os_, kw)
    env.close()
