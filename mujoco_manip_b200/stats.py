"""Episode statistics across GPUs.

Envs shard over ranks with no data-path exchange; the only collective of the whole system is this
optional reduction of the per-rank statistics vector [episodes, successes, sum length, sum return,
non-finite resets, overflow episodes, failed placements, spare] that mm_post_step / mm_sample_episode accumulate
on the device (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist

STAT_NAMES = ("episodes", "successes", "sum_length", "sum_return", "nonfinite_resets", "overflow_episodes",
              "failed_placements", "spare")


def shard_range(num_envs_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of global env ids owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(num_envs_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_stats(local: torch.Tensor, group=None) -> torch.Tensor:
    """[world, 8] statistics of every rank (all_gather); the local vector when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local.reshape(1, -1).clone()
    out = [torch.empty_like(local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, local.contiguous(), group=group)
    return torch.stack(out)


def summarize(stats: torch.Tensor) -> dict:
    """Whole-job summary from the [world, 8] (or [8]) statistics."""
    s = stats.reshape(-1, len(STAT_NAMES)).sum(dim=0).double()
    ep = max(float(s[0]), 1.0)
    return {"episodes": float(s[0]), "success_rate": float(s[1]) / ep, "mean_length": float(s[2]) / ep,
            "mean_return": float(s[3]) / ep, "nonfinite_resets": float(s[4]), "overflow_episodes": float(s[5]),
            "failed_placements": float(s[6])}
