"""Episode statistics across ranks -- the ONLY collective of the system.

Envs shard across GPUs with no exchange inside ``step``/``reset``; at log cadence each rank contributes the
nine device-side accumulators of ``rd_read_stats`` (episodes, return_sum, progress_sum, ...) and every rank
receives the per-rank table and its sum.  Mirrors what ``tools.simulate`` aggregates per episode on the host
[REF dreamer/tools.py:159-206] and what ``callbacks.summarize_episode`` logs [REF dreamer/callbacks.py:74-100].
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

STAT_KEYS = ("episodes", "return_sum", "progress_sum", "length_sum", "collisions", "laps_completed", "env_steps",
             "timeouts", "max_progress_sum")


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous env-id range [lo, hi) owned by ``rank``; remainders go to the lowest ranks."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_stats(stats: Dict[str, float], device=None, group=None) -> Tuple[Dict[str, float], List[Dict[str, float]]]:
    """all_gather of one rank's ``rd_read_stats`` dict.  Returns (sum over ranks, per-rank list).
    Works with NCCL (device = this rank's GPU) and gloo (device = cpu); a no-op outside a process group."""
    vec = torch.tensor([float(stats[k]) for k in STAT_KEYS], dtype=torch.float64, device=device or "cpu")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(zip(STAT_KEYS, vec.tolist())), [dict(zip(STAT_KEYS, vec.tolist()))]
    world = dist.get_world_size(group)
    flat = torch.empty(world * len(STAT_KEYS), dtype=torch.float64, device=vec.device)
    dist.all_gather_into_tensor(flat, vec, group=group)   # 1-D output: accepted by both NCCL and gloo
    table = flat.view(world, len(STAT_KEYS))
    per_rank = [dict(zip(STAT_KEYS, row.tolist())) for row in table]
    return dict(zip(STAT_KEYS, table.sum(0).tolist())), per_rank


def summarize(total: Dict[str, float]) -> Dict[str, float]:
    """Means the reference logs per episode: return, progress (laps), length [REF dreamer/callbacks.py:80-92]."""
    n = max(total.get("episodes", 0.0), 1.0)
    return {"episodes": total.get("episodes", 0.0), "mean_return": total["return_sum"] / n,
            "mean_progress": total["progress_sum"] / n, "mean_length": total["length_sum"] / n,
            # tools.simulate's per-episode statistic: max over the episode of lap + progress - 1 [REF dreamer/tools.py:181,195]
            "mean_max_progress": total.get("max_progress_sum", 0.0) / n,
            "collision_rate": total["collisions"] / n, "timeout_rate": total["timeouts"] / n,
            "env_steps": total["env_steps"]}
