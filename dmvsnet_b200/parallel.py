"""Multi-GPU plumbing for the cost-volume path: one process per GPU, ``torch.distributed`` (NCCL over NVLink).

What shards and what does not (DESIGN.md "Multi-GPU"):

* **views** - every reference view is an independent forward.  Throughput scaling = replicas, no collective
  (``view_shard``).  This is what ``bench.py --gpus N`` measures.
* **W1 depth planes** - cost-volume planes are independent given the (replicated) features, so one view's
  warp+corr can be split over ranks by plane range and re-assembled with a single all-gather
  (``gather_planes`` / ``warp_corr_depth_sharded``).  Exact: every plane is computed by exactly one rank with the
  same kernel, so sharded == unsharded bit for bit.
* **R1 regularisation nets** - do NOT shard by depth (the receptive field spans every plane).  Their two branches
  (cosR_small / cosR_huge, reference module.py:343-349) are independent networks on the same input, so two ranks take one
  each and exchange their two logit channels with one all-gather (``regnet_branch_sharded``): exact, and the single-view
  latency mode of ``MVSNet.cascade(..., branch_group=...)``.

  For more than two ranks they shard by ROWS with a halo: ``row_bands`` plans bands whose edges are multiples of 8 (three
  stride-2 levels) with a 32-row halo - the U-Nets' receptive field is +-30 rows (1+1+2+2+4+4+8 down, 4+2+1 up, 1 for prob);
  measured on the CPU restatement of the nets, a 24-row halo leaves a 2.5e-3 error, 32 rows reproduce the unsharded logits
  bit for bit - and ``gather_rows`` re-assembles the owned rows with one all-gather.  Host logic only so far (gloo-tested in
  tests/test_parallel_gloo.py with a CPU regularisation net as the per-band operator); the device side needs a row offset in W1 / S1 / E2, where the absolute row
  enters the homography and the checkerboard.

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


def exchange_channel_halves(logits: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """``logits`` [B,4,D,h,w]: rank r (of 2) has written channels 2r, 2r+1; after the call every rank holds all four.
    B == 1: the halves are contiguous, one in-place all-gather without staging copies; otherwise via contiguous copies."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world != 2:
        raise ValueError("branch sharding needs a 2-rank group, got %d ranks" % world)
    b = logits.shape[0]
    if b == 1 and logits.is_contiguous() and logits.is_cuda:
        flat = logits.view(2, -1)
        dist.all_gather_into_tensor(flat, flat[rank], group=group)  # NCCL in-place form: input = output + rank * count
        return logits
    mine = logits[:, 2 * rank:2 * rank + 2].contiguous()
    parts = [torch.empty_like(mine) for _ in range(2)]
    dist.all_gather(parts, mine, group=group)
    for r in range(2):
        logits[:, 2 * r:2 * r + 2] = parts[r]
    return logits


def regnet_branch_sharded(pack, cost: Optional[torch.Tensor], cost_cells: Optional[torch.Tensor],
                          group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """CostRegNet / CostRegNet_refine over 2 ranks: rank r runs branch r (dmvs_regnet_forward_branches_f32), one all-gather of
    the logit halves.  Bit-identical to the unsharded call except for conv0, which the single-GPU path runs as one paired
    launch: same arithmetic per output channel, same bits."""
    from . import ops
    rank = dist.get_rank(group)
    src = cost if cost is not None else cost_cells
    if cost is not None:
        b, _, d, h, w = cost.shape
    else:
        b, d, h, w = cost_cells.shape[0], cost_cells.shape[1], cost_cells.shape[2], cost_cells.shape[3] - 1
    logits = torch.empty(b, 4, d, h, w, device=src.device, dtype=torch.float32)
    ops.regnet_forward(pack, cost, cost_cells=cost_cells, branch_mask=1 << rank, out=logits)
    return exchange_channel_halves(logits, group)



# ----------------------------------------------------------------------------- row bands (R1 over more than two ranks)
REGNET_HALO_ROWS = 32  # receptive field of CostRegNet / CostRegNet_refine along H and W is +-30, bands are cut at multiples of 8


def row_bands(num_rows: int, world: int, halo: int = REGNET_HALO_ROWS, multiple: int = 8) -> List[Tuple[int, int, int, int]]:
    """Per rank ``(own_lo, own_hi, band_lo, band_hi)``: the rows a rank owns (contiguous, edges at multiples of ``multiple``,
    balanced to within one block) and the rows it has to compute on, i.e. its own rows plus ``halo`` rows either side clipped
    to the image.  Ranks beyond the number of blocks own nothing (empty range, empty band)."""
    if num_rows % multiple:
        raise ValueError("num_rows=%d must be a multiple of %d" % (num_rows, multiple))
    out = []
    for lo, hi in plane_shards(num_rows // multiple, world):
        own_lo, own_hi = lo * multiple, hi * multiple
        if own_hi == own_lo:
            out.append((own_lo, own_hi, own_lo, own_hi))
        else:
            out.append((own_lo, own_hi, max(own_lo - halo, 0), min(own_hi + halo, num_rows)))
    return out


def gather_rows(compute: Callable[[int, int, int, int], Optional[torch.Tensor]], num_rows: int,
                group: Optional[dist.ProcessGroup] = None, row_dim: int = -2, halo: int = REGNET_HALO_ROWS,
                multiple: int = 8, like: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Each rank calls ``compute(own_lo, own_hi, band_lo, band_hi)`` and returns the result for its OWN rows (the halo rows it
    computed on are its business); one all-gather re-assembles all ``num_rows`` along ``row_dim`` on every rank.  A rank that
    owns nothing returns None from ``compute`` and passes ``like`` (any tensor with the result's dtype / device and its shape
    except along ``row_dim``)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bands = row_bands(num_rows, world, halo, multiple)
    own_lo, own_hi, band_lo, band_hi = bands[rank]
    mine = compute(own_lo, own_hi, band_lo, band_hi) if own_hi > own_lo else None
    proto = mine if mine is not None else like
    if proto is None:
        raise ValueError("rank %d owns no rows and got no `like` tensor to shape its part of the collective" % rank)
    dim = row_dim % proto.dim()
    if mine is not None and mine.shape[dim] != own_hi - own_lo:
        raise ValueError("compute returned %d rows for the range [%d, %d)" % (mine.shape[dim], own_lo, own_hi))
    widest = max(b[1] - b[0] for b in bands)
    rest = tuple(proto.shape[:dim]) + tuple(proto.shape[dim + 1:])
    send = proto.new_zeros((widest,) + rest)
    if mine is not None:
        send[: own_hi - own_lo] = mine.movedim(dim, 0)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    parts = [recv[r][: bands[r][1] - bands[r][0]] for r in range(world)]
    return torch.cat(parts, 0).movedim(0, dim).contiguous()
