"""world_size-2 gloo test of the multi-GPU host logic: env sharding + the episode-statistics gather
(the only collective of the system).  The env step itself needs a GPU; here each rank steps its shard
of the CPU oracle so the gathered totals can be checked against a single-process run of the whole batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from racing_dreamer_b200.stats import STAT_KEYS, gather_stats, shard_range, summarize

N_TOTAL, STEPS = 96, 12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_stats(lo, hi):
    """Steps envs [lo, hi) of one logical batch (global ids via env_id_offset) on the CPU oracle."""
    from oracle import Oracle, default_config
    from racing_dreamer_b200 import _abi, load_track
    cfg = default_config()
    cfg.n_envs = hi - lo
    cfg.env_id_offset = lo
    cfg.action_repeat = 8
    cfg.auto_reset = 1
    cfg.reset_mode = _abi.RESET_RANDOM
    cfg.seed = 5
    cfg.time_limit_steps = 6
    orc = Oracle(cfg, [load_track("treitlstrasse_v2")], n_threads=1)
    orc.reset(mode=_abi.RESET_RANDOM)
    acts = np.random.RandomState(3).uniform(-1, 1, (STEPS, N_TOTAL, 2)).astype(np.float32)
    for k in range(STEPS):
        orc.step(acts[k, lo:hi])
    return orc.read_stats()


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(N_TOTAL, rank, world)
        total, per_rank = gather_stats(_oracle_stats(lo, hi))
        assert len(per_rank) == world
        torch.save({"total": total, "per_rank": per_rank, "range": (lo, hi)}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    for n, w in ((96, 2), (10, 4), (7, 8), (1048576, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_gather_stats_world2_matches_single_process(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert res[0]["total"] == res[1]["total"]                      # every rank sees the same table
    assert res[0]["range"] == (0, 48) and res[1]["range"] == (48, 96)
    whole = _oracle_stats(0, N_TOTAL)                              # sharding must not change any episode
    for k in STAT_KEYS:
        assert abs(res[0]["total"][k] - whole[k]) <= 1e-9 * max(1.0, abs(whole[k])), k
    assert res[0]["total"]["env_steps"] == N_TOTAL * STEPS
    s = summarize(res[0]["total"])
    assert s["episodes"] > 0 and np.isfinite(s["mean_return"])


def test_gather_stats_without_process_group():
    total, per_rank = gather_stats({k: float(i) for i, k in enumerate(STAT_KEYS)})
    assert total["timeouts"] == 7.0 and len(per_rank) == 1
