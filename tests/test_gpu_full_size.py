"""BASELINE.json's configurations at FULL size on the GPU, checked through size-independent properties (the oracle cannot
step 65536+ envs in test time) and against the oracle on sub-batches.

The law that carries the parity of the small cases to the big ones: an env's trajectory depends only on its GLOBAL env id
(reset sampling, noise counters), its map and its actions -- not on how many envs share the launch, on the CTA/warp it
lands in, or on the chunking of the host-facing path.  So envs [off, off + m) of a full-size batch must reproduce, bit for
bit, an m-env batch created with env_id_offset = off, and that small batch is the one held against the oracle.
"""
import os

import numpy as np
import pytest

from racing_dreamer_b200 import _abi

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1
LIDAR_TOL_M = 1e-3


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


def make_env(**kw):
    from racing_dreamer_b200 import BatchedRaceEnv, EnvConfig
    return BatchedRaceEnv(EnvConfig(**kw), device="cuda:0")


def _actions(rng, n, kind):
    if kind == "random":
        return rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    a = np.empty((n, 2), np.float32)
    a[:, 0] = 0.6
    a[:, 1] = 0.8 * np.sin(rng.uniform(0, 2 * np.pi, n))
    return a


def _run_sub_batch_identity(torch, n, m, off, steps, oracle_steps, actions="scripted", **cfg):
    """big: n envs; small: envs [off, off + m) of it as their own batch; oracle: the small batch on the CPU."""
    from oracle import Oracle
    big = make_env(n_envs=n, **cfg)
    sub_ids = None
    if cfg.get("map_ids") is None and len(cfg["tracks"]) > 1:
        A = cfg.get("agents_per_world", 1)
        sub_ids = ((np.arange(off, off + m) // A) % len(cfg["tracks"])).astype(np.int32)
    small = make_env(n_envs=m, env_id_offset=off, map_ids=sub_ids, **cfg)
    orc = Oracle(small.cfg, small.tracks, small.map_ids, n_threads=THREADS)
    ob, os_ = big.reset(), small.reset()
    orc.reset(mode=int(small.cfg.reset_mode))
    assert torch.equal(ob["lidar"][off:off + m], os_["lidar"])
    rng = np.random.RandomState(n % 1000)
    dones = 0
    for k in range(steps):
        a = _actions(rng, n, actions)
        ta = torch.from_numpy(a).cuda()
        obs_b, rew_b, done_b, info_b = big.step(ta)
        obs_s, rew_s, done_s, info_s = small.step(ta[off:off + m].contiguous())
        for key in ("lidar", "pose", "velocity", "speed"):
            assert torch.equal(obs_b[key][off:off + m], obs_s[key]), (key, k)
        if "lidar_occupancy" in obs_b:
            assert torch.equal(obs_b["lidar_occupancy"][off:off + m], obs_s["lidar_occupancy"]), k
        assert torch.equal(rew_b[off:off + m], rew_s) and torch.equal(done_b[off:off + m], done_s)
        for key in ("progress", "lap", "time", "flags", "rank", "opponent_collisions"):
            assert torch.equal(info_b[key][off:off + m], info_s[key]), (key, k)
        # laws that need no reference: ranges inside the sensor's interval, finite rewards, laps never below 1
        lo, hi = float(big.cfg.lidar_range_min), float(big.cfg.lidar_range_max)
        assert float(obs_b["lidar"].min()) >= np.float32(lo) and float(obs_b["lidar"].max()) <= np.float32(hi)
        assert bool(torch.isfinite(rew_b).all()) and int(info_b["lap"].min()) >= 1
        dones += int(done_b.sum())
        if k < oracle_steps:                      # the small batch against the oracle
            ref = orc.step(a[off:off + m])
            assert np.array_equal(done_s.cpu().numpy().astype(np.uint8), ref["done"]), k
            assert np.array_equal(info_s["flags"].cpu().numpy(), ref["flags"]), k
            assert np.array_equal(info_s["lap"].cpu().numpy(), ref["lap"])
            assert np.abs(obs_s["lidar"].cpu().numpy() - ref["lidar"]).max() <= LIDAR_TOL_M
            if "lidar_occupancy" in obs_s:
                assert np.array_equal(obs_s["lidar_occupancy"].cpu().numpy()[..., 0], ref["occupancy"]), k
    fb, ib = big.get_state()
    fs, is_ = small.get_state()
    assert torch.equal(fb[:, off:off + m], fs) and torch.equal(ib[:, off:off + m], is_)
    st = big.read_stats()
    A = max(1, int(big.cfg.agents_per_world))
    assert st["env_steps"] == n * steps
    if A == 1:
        assert st["episodes"] == dones            # every done is one finished episode (auto-reset)
    big.close(); small.close()
    return dones


def test_config4_treitlstrasse_65536_envs_random_actions(torch_cuda):
    """BASELINE config 4: collisions / laps / time limits fire, auto-reset; flags verified on a 4096-env sub-batch."""
    dones = _run_sub_batch_identity(torch_cuda, 65536, 4096, 40960, steps=30, oracle_steps=12, actions="random",
                                    tracks=("treitlstrasse_v2",), action_repeat=8, auto_reset=True, reset_mode="random",
                                    seed=4, laps=1, time_limit_steps=2000 // 8)
    assert dones > 65536                          # random actions: every env crashes a few times in 30 steps


def test_config5_mixed_maps_131072_envs(torch_cuda):
    """BASELINE config 5, one GPU's share: Barcelona/Austria alternating by env index."""
    _run_sub_batch_identity(torch_cuda, 131072, 2048, 100352, steps=10, oracle_steps=5,
                            tracks=("barcelona", "austria"), action_repeat=8, auto_reset=True, reset_mode="random", seed=5,
                            time_limit_steps=2000 // 8)


def test_config3_columbia_16384_envs_occupancy(torch_cuda):
    """BASELINE config 3: lidar_occupancy at full size; a 256-env sub-batch bit-exact against the oracle's images."""
    _run_sub_batch_identity(torch_cuda, 16384, 256, 9984, steps=6, oracle_steps=6, tracks=("columbia",), action_repeat=8,
                            obs_type="lidar_occupancy", auto_reset=True, reset_mode="random", seed=3,
                            time_limit_steps=2000 // 8)


def test_worlds_of_four_cars_at_scale(torch_cuda):
    """16384 worlds x 4 cars: the sub-batch law holds for whole worlds; every world resets as one."""
    _run_sub_batch_identity(torch_cuda, 65536, 1024, 20480, steps=12, oracle_steps=6, actions="random",
                            tracks=("austria",), action_repeat=4,
                            agents_per_world=4, agent_tasks=("maximize_progress",) + ("n_step_progress",) * 3,
                            auto_reset=True, reset_mode="random_ball", ball_spacing=0.8, seed=7, time_limit_steps=40)


def test_host_facing_step_at_config2_size_equals_device_step(torch_cuda):
    """The e2e call of bench.py (rd_step_host, 8 chunks) at config-2 size == rd_step on the same envs, bit for bit."""
    torch = torch_cuda
    from racing_dreamer_b200 import BatchedRaceEnv, EnvConfig
    from racing_dreamer_b200.host import HostSteppedEnv
    ec = EnvConfig(tracks=("austria",), n_envs=4096, action_repeat=8, auto_reset=True, reset_mode="random", seed=1,
                   time_limit_steps=250)
    dev, host = BatchedRaceEnv(ec, device="cuda:0"), HostSteppedEnv(ec, device="cuda:0", n_shards=8)
    dev.reset(); host.reset()
    rng = np.random.RandomState(0)
    for k in range(20):
        a = _actions(rng, 4096, "scripted")
        obs, rew, done, info = dev.step(torch.from_numpy(a).cuda())
        out = host.step(a)
        assert np.array_equal(out["lidar"], obs["lidar"].cpu().numpy()), k
        assert np.array_equal(out["reward"], rew.cpu().numpy()) and np.array_equal(out["flags"], info["flags"].cpu().numpy())
    dev.close(); host.close()


def test_two_groups_async_host_steps_equal_the_synchronous_call(torch_cuda):
    """step_async/step_wait (rd_step_host_begin/_end) on two half-batches in flight == step() on each, bit for bit; a
    second begin without end, or a reset while a step is pending, is refused."""
    import dataclasses
    from racing_dreamer_b200 import EnvConfig
    from racing_dreamer_b200.host import HostSteppedEnv
    half = 1024
    base = EnvConfig(tracks=("austria",), n_envs=half, action_repeat=8, auto_reset=True, reset_mode="random", seed=2,
                     time_limit_steps=30)
    cfgs = [dataclasses.replace(base, env_id_offset=g * half) for g in range(2)]
    sync = [HostSteppedEnv(c, device="cuda:0", n_shards=4) for c in cfgs]
    asyn = [HostSteppedEnv(c, device="cuda:0", n_shards=4) for c in cfgs]
    for e in sync + asyn:
        e.reset()
    rng = np.random.RandomState(3)
    acts = rng.uniform(-1, 1, (40, 2, half, 2)).astype(np.float32)
    asyn[0].step_async(acts[0, 0])
    with pytest.raises(RuntimeError, match="pending"):
        asyn[0].step_async(acts[0, 0])
    with pytest.raises(RuntimeError, match="pending"):
        asyn[0].reset()
    for k in range(40):
        asyn[1].step_async(acts[k, 1])
        want0 = {key: v.copy() for key, v in sync[0].step(acts[k, 0]).items()}
        got0 = asyn[0].step_wait()
        for key in want0:
            assert np.array_equal(got0[key], want0[key]), (key, k)
        if k + 1 < 40:
            asyn[0].step_async(acts[k + 1, 0])
        want1 = sync[1].step(acts[k, 1])
        got1 = asyn[1].step_wait()
        for key in want1:
            assert np.array_equal(got1[key], want1[key]), (key, k)
    with pytest.raises(RuntimeError, match="pending"):
        asyn[1].step_wait()
    for e in sync + asyn:
        e.close()
