"""GPU parity of the on-device follow-the-gap policy, through the C ABI (rd_policy_gap_follower,
rd_rollout_gap_follower): against the golden commands of the unmodified reference node, against the numpy restatement
in closed loop, and through laps-without-collision properties at batch sizes the oracle could not replay.

Tolerance: 1e-9 on steering angle / speed / heading (float64 on both sides; the only differences are the summation
order of two means and acos vs libm), 1e-6 on the float32 actions."""
import os

import numpy as np
import pytest

from racing_dreamer_b200 import _abi
from oracle.gap_follower import GapFollowerOracle, GapFollowerParams
from test_cpu_gap_follower import golden_sequences

pytestmark = pytest.mark.gpu
TOL = 1e-9
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


def env_lidar_from_arc(arc, s0, s1, nb):
    ros = np.zeros(nb, np.float32)
    ros[s0:s1 + 1] = arc
    return ros[::-1].copy()          # env order: index 0 = left


def expected_action(cfg, sa, sp, v, speed_scale=0.5, gain=1.0):
    veh = cfg.vehicle
    vt = sp * speed_scale
    motor = min(max(vt * veh.c_drag / veh.a_drive + gain * (vt - v), cfg.action_low[0]), cfg.action_high[0])
    steer = min(max(sa / (veh.steer_gain * veh.steer_max), cfg.action_low[1]), cfg.action_high[1])
    if cfg.rescale_actions:
        motor = (motor - cfg.action_low[0]) / (cfg.action_high[0] - cfg.action_low[0]) * 2.0 - 1.0
        steer = (steer - cfg.action_low[1]) / (cfg.action_high[1] - cfg.action_low[1]) * 2.0 - 1.0
    return np.float32(motor), np.float32(steer)


def test_policy_matches_reference_golden(torch_cuda, golden_dir):
    """teacher-forced scans: env 0 replays the golden sequence, envs 1..3 the same sequence from shifted starts"""
    torch = torch_cuda
    from racing_dreamer_b200 import GapFollowerPolicy
    for si, R, s0, s1, nb, arcs, cmds in golden_sequences(golden_dir):
        n = 4
        env = make_env(tracks=("austria",), n_envs=n, action_repeat=R, n_beams=nb)
        pol = GapFollowerPolicy(env)
        assert (pol.params.arc_first, pol.params.arc_last) == (s0, s1)
        oracles = [GapFollowerOracle(GapFollowerParams(n_beams=nb, dt=R * 0.01)) for _ in range(n)]
        speeds = np.linspace(0.0, 3.0, n).astype(np.float32)
        T = len(arcs)
        for k in range(T):
            rows = np.stack([env_lidar_from_arc(arcs[(k + 7 * j) % T], s0, s1, nb) for j in range(n)])
            act = pol.act(torch.from_numpy(rows).cuda(), speed=torch.from_numpy(speeds).cuda(), debug=True).cpu().numpy()
            dbg = pol.debug.cpu().numpy()
            for j in range(n):
                pub, sa, sp, hd = oracles[j](rows[j][::-1].astype(np.float64))
                assert abs(dbg[j, 0] - sa) <= TOL and abs(dbg[j, 1] - sp) <= TOL and abs(dbg[j, 2] - hd) <= TOL, (si, k, j)
                m, s = expected_action(env.cfg, sa, sp, float(speeds[j]))
                assert abs(act[j, 0] - m) <= 1e-6 and abs(act[j, 1] - s) <= 1e-6
            # env 0 is the recorded sequence itself: compare with what the reference node published
            assert abs(dbg[0, 0] - cmds[k, 1]) <= TOL and abs(dbg[0, 1] - cmds[k, 2]) <= TOL and abs(dbg[0, 2] - cmds[k, 3]) <= TOL
        env.close()


@pytest.mark.parametrize("track,R,nb", [("austria", 4, 1080), ("treitlstrasse_v2", 8, 1080), ("columbia", 4, 2048),
                                        ("austria", 4, 1440)])
def test_closed_loop_vs_oracle(torch_cuda, track, R, nb):
    """GPU env + GPU policy; the numpy controller sees the GPU's scans, the CPU oracle env is stepped with the GPU's
    actions: controller outputs within TOL every step, env results at the usual parity bars, resets clear the PID."""
    torch = torch_cuda
    from racing_dreamer_b200 import GapFollowerPolicy
    from oracle import Oracle
    n = 48
    env = make_env(tracks=(track,), n_envs=n, action_repeat=R, auto_reset=True, reset_mode="random", seed=21,
                   time_limit_steps=25, n_beams=nb)
    orc = Oracle(env.cfg, env.tracks, env.map_ids, n_threads=THREADS)
    pol = GapFollowerPolicy(env)
    ctl = [GapFollowerOracle(GapFollowerParams(n_beams=nb, dt=R * 0.01)) for _ in range(n)]
    obs = env.reset()
    orc.reset(mode=int(env.cfg.reset_mode))
    resets = 0
    for k in range(70):
        lidar = obs["lidar"].cpu().numpy()
        act = pol.act(debug=True).cpu().numpy().copy()
        dbg = pol.debug.cpu().numpy()
        v = orc.f64[_abi.S_V].copy()
        for j in range(n):
            pub, sa, sp, hd = ctl[j](lidar[j][::-1].astype(np.float64))
            assert abs(dbg[j, 0] - sa) <= TOL and abs(dbg[j, 1] - sp) <= TOL and abs(dbg[j, 2] - hd) <= TOL, (k, j)
            m, s = expected_action(env.cfg, sa, sp, v[j])
            assert abs(act[j, 0] - m) <= 1e-6 and abs(act[j, 1] - s) <= 1e-6
        obs, rew, done, info = env.step(pol.actions)
        ref = orc.step(act)
        d = done.cpu().numpy()
        assert np.array_equal(d.astype(np.uint8), ref["done"])
        assert np.abs(obs["lidar"].cpu().numpy() - ref["lidar"]).max() <= 1e-3
        assert np.allclose(rew.cpu().numpy(), ref["reward"], rtol=1e-5, atol=1e-6)
        for j in np.nonzero(d)[0]:
            ctl[j].reset()          # the env kernel cleared the device-side controller of an auto-reset env
            resets += 1
    assert resets >= n              # the 25-step time limit guarantees every controller was cleared at least once
    env.close()


def test_rollout_equals_stepwise(torch_cuda):
    """rd_rollout_gap_follower (no host in the loop) == act() + step() called from the host, bit for bit"""
    torch = torch_cuda
    from racing_dreamer_b200 import GapFollowerPolicy
    outs = []
    for mode in ("rollout", "stepwise"):
        env = make_env(tracks=("columbia",), n_envs=256, action_repeat=4, auto_reset=True, reset_mode="random", seed=5)
        pol = GapFollowerPolicy(env)
        env.reset()
        if mode == "rollout":
            pol.rollout(40)
        else:
            for _ in range(40):
                env.step(pol.act())
        f, i = env.get_state()
        outs.append((f.cpu().numpy(), i.cpu().numpy(), env.buf["lidar"].cpu().numpy().copy(), env.read_stats()))
        env.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2]) and outs[0][3] == outs[1][3]


@pytest.mark.parametrize("track", ["austria", "columbia", "treitlstrasse_v2", "barcelona"])
def test_gap_follower_laps_without_collision(torch_cuda, track):
    """size-independent property: from the grid start the controller drives every env of a large batch around the
    track -- progress only ever increases, nobody touches a wall, and laps are counted"""
    torch = torch_cuda
    from racing_dreamer_b200 import GapFollowerPolicy
    n = 4096
    env = make_env(tracks=(track,), n_envs=n, action_repeat=4, auto_reset=False, reset_mode="grid", laps=50)
    pol = GapFollowerPolicy(env)
    env.reset()
    last = torch.zeros(n, device="cuda")
    for chunk in range(10):
        obs, rew, done, info = pol.rollout(200)
        total = info["lap"].float() - 1.0 + info["progress"]
        assert not bool(done.any()) and not bool(info["wall_collision"].any()) and not bool(info["wrong_way"].any())
        assert bool((total >= last - 1e-6).all())
        last = total.clone()
    # identical envs from the same start stay identical (determinism across the batch)
    assert float(total.max() - total.min()) == 0.0
    assert float(total[0]) > (0.4 if track == "barcelona" else 0.9)
    env.close()


def test_policies_in_worlds_of_cars(torch_cuda):
    """On-device policies with agents_per_world > 1: one controller / one latent per CAR, cleared when the car's world is
    reset; the no-host rollout equals the host-driven loop bit for bit; the cars really interact (scans differ from the
    single-car ones, contacts end worlds)."""
    torch = torch_cuda
    from racing_dreamer_b200 import DreamerPolicy, GapFollowerPolicy
    kw = dict(tracks=("austria",), n_envs=512, agents_per_world=4, action_repeat=4, auto_reset=True, reset_mode="random_ball",
              ball_spacing=0.9, seed=5, time_limit_steps=60)
    for make_policy in (lambda e: GapFollowerPolicy(e), lambda e: DreamerPolicy(e, "austria_dreamer", noise="philox")):
        outs = []
        for mode in ("rollout", "stepwise"):
            env = make_env(**kw)
            pol = make_policy(env)
            env.reset()
            if mode == "rollout":
                pol.rollout(50)
            else:
                for _ in range(50):
                    env.step(pol.act())
            f, i = env.get_state()
            outs.append((f.cpu().numpy(), i.cpu().numpy(), env.buf["lidar"].cpu().numpy().copy(), env.read_stats()))
            env.close()
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        assert np.array_equal(outs[0][2], outs[1][2])
        for key in outs[0][3]:      # the statistics are float64 atomics: the order of the additions is not fixed
            assert abs(outs[0][3][key] - outs[1][3][key]) <= 1e-9 * max(1.0, abs(outs[1][3][key])), key
        st = outs[0][3]
        assert st["episodes"] > 0 and st["env_steps"] == 512 * 50
        ep = outs[0][1][_abi.I_EPISODE].reshape(-1, 4)
        assert np.all(ep == ep[:, :1])          # worlds reset as a whole
