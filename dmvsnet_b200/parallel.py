"""Multi-GPU plumbing for the cost-volume path: one process per GPU, ``torch.distributed`` (NCCL over NVLink).

What shards and what does not (DESIGN.md "Multi-GPU"):

* **views** - every reference view is an independent forward.  Throughput scaling = replicas, no collective
  (``view_shard``).  This is what ``bench.py --gpus N`` measures.
* **W1 depth planes** - cost-volume planes are independent given the (replicated) features, so one view's
  warp+corr can be split over ranks by plane range and re-assembled with a single all-gather
  (``gather_planes`` / ``warp_corr_depth_sharded``).  Exact: every plane is computed by exactly one rank with the
  same kernel, so sharded == unsharded bit for bit.
* **R1 regularisation nets** - do NOT shard by depth (receptive field +-24 planes >= D) - replicated.

The gather logic is device agnostic (it is exercised with gloo on CPU in tests/test_parallel_gloo.py); only
``warp_corr_depth_sharded`` touches the CUDA op.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def plane_shards(num_planes: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [lo, hi) plane ranges; ranks beyond ``num_planes`` get empty ranges."""
    base, extra = divmod(num_planes, world)
    out, lo = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def view_shard(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of independent reference views to ranks (replica mode)."""
    return list(range(rank, num_views, world))


def gather_planes(compute: Callable[[int, int], torch.Tensor], num_planes: int, group: Optional[dist.ProcessGroup] = None,
                  plane_dim: int = 2) -> torch.Tensor:
    """Each rank computes its plane range with ``compute(lo, hi)`` (-> tensor with hi-lo planes along ``plane_dim``);
    one all-gather re-assembles the full volume on every rank.  Uneven shards are padded to the largest shard."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    shards = plane_shards(num_planes, world)
    lo, hi = shards[rank]
    widest = max(b - a for a, b in shards)
    mine = compute(lo, hi) if hi > lo else None
    ref = mine
    if ref is None:  # this rank has no planes: it still takes part in the collective with a dummy of the right shape
        probe = compute(0, 1)
        ref = probe
        mine = probe.narrow(plane_dim, 0, 0)
    mine = mine.movedim(plane_dim, 0).contiguous()            # [planes, ...]
    pad_shape = (widest,) + tuple(mine.shape[1:])
    send = mine.new_zeros(pad_shape)
    send[: mine.shape[0]] = mine
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    parts = [recv[r][: shards[r][1] - shards[r][0]] for r in range(world)]
    return torch.cat(parts, 0).movedim(0, plane_dim).contiguous()


def warp_corr_depth_sharded(features: Sequence[torch.Tensor], rt: torch.Tensor, hyp: torch.Tensor,
                            group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """W1 for one view split over the ranks by depth plane; features / hypotheses replicated on every rank."""
    from . import ops
    b, d, h, w = hyp.shape

    def compute(lo: int, hi: int) -> torch.Tensor:
        full = torch.empty(b, 2, d, h, w, device=hyp.device, dtype=torch.float32)
        ops.warp_corr(features, rt, hyp, d_range=(lo, hi), out=full)
        return full[:, :, lo:hi]

    return gather_planes(compute, d, group, plane_dim=2)
