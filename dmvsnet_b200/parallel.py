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


# ----------------------------------------------------------------------------- single-view mode: the whole forward over R ranks
# One reference view on R GPUs (BASELINE config 4: T&T 1920x1056, N = 11 on 8 x B200).  What is exchanged, per view:
#   * FeatureNet shards by VIEW (reference mvsnet.py:199-202 loops over views): every rank computes the reference view (the only
#     map W1 reads in fp32, and only its own rows of it) and its share of the source views, then ONE all-gather per feature map of
#     the fp16 source maps (the format W1's staged kernel gathers from) makes all sources resident everywhere - stage-1 maps
#     first, the later stages' gathers overlap the earlier stages' compute on NCCL's stream;
#   * the stage loop shards by ROW BANDS with a 32-row halo (``row_bands``): W1 is pointwise in the reference pixel, so a rank
#     computes the cost volume of its band + halo itself from the resident features; the U-Nets see real data 32 rows beyond the
#     rows the rank owns (their receptive field is +-30), so the owned rows carry the bits of the unsharded result;
#   * per stage two small exchanges re-assemble per-pixel maps: ``depth_values_c`` [B,4,h,w] after the main pass (the refine
#     pass needs it 32 rows beyond the owned rows) and ``depth`` / confidence [B,h,w] after the refine pass (the next stage's
#     sampler upsamples it: +-1 row).  They are all-gathers of the owned rows (equal slots of the widest band).
# Nothing else crosses NVLink: no cost volume, no logits, no probability volume (those stay row-sharded).


class _CudaBackend:
    """The per-band operators of ``cascade_row_sharded`` on the CUDA library (a CPU stand-in with the same methods drives the
    gloo test)."""

    def __init__(self, net):
        self.net = net

    def hypotheses_first(self, dv, ndepth, shape, inverse):
        from . import ops
        return ops.hypotheses_first(dv, ndepth, shape, inverse)

    def hypotheses_next(self, last, ndepth, ip, shape, inverse):
        from . import ops
        return ops.hypotheses_next(last, ndepth, ip, shape, inverse)

    def regularised_logits(self, key, ref_band, srcs, rt, hyp_band, row0, stage, refine):
        """W1 on a row band (fp16-staged kernel, conv0 cells) + the stage's regularisation net -> logits [B,4,D,rows,w]."""
        from . import ops
        _, cells = ops.warp_corr([ref_band] + list(srcs), rt, hyp_band, want_f32=False, want_cells=True, layout="h16", row0=row0)
        mod = (self.net.cost_regularization_refine if refine else self.net.cost_regularization)[stage]
        return mod(None, cost_cells=cells)

    def depth_head(self, logits, hyp, interval):
        from . import ops
        return ops.depth_head(logits, hyp, interval, want_prob=False)[1:]

    def refine_head(self, logits, hyp_c, interval):
        from . import ops
        return ops.refine_head(logits, hyp_c, interval, 5.0)


def _assemble_rows(part: torch.Tensor, own: Tuple[int, int], num_rows: int, group) -> torch.Tensor:
    """``part``: this rank's OWNED rows along dim -2 ([B,C,own_hi - own_lo,w]) -> the full map [B,C,num_rows,w] on every rank.
    One all-gather of equal slots (the widest band, zero padded): every row crosses NVLink once."""
    world = dist.get_world_size(group)
    bands = row_bands(num_rows, world)
    widest = max(b[1] - b[0] for b in bands)
    send = part.new_zeros(tuple(part.shape[:-2]) + (widest, part.shape[-1]))
    if own[1] > own[0]:
        send[..., : own[1] - own[0], :] = part
    recv = part.new_empty((world,) + tuple(send.shape))
    if send.is_cuda:
        dist.all_gather_into_tensor(recv, send, group=group)
    else:  # gloo (the CPU tests) has no flat all-gather
        dist.all_gather(list(recv.unbind(0)), send, group=group)
    return torch.cat([recv[r][..., : bands[r][1] - bands[r][0], :] for r in range(world) if bands[r][1] > bands[r][0]], dim=-2)


def cascade_row_sharded(net, ref_feats, src_feats, proj_matrices, depth_values: torch.Tensor, image_hw: Sequence[int],
                        group: Optional[dist.ProcessGroup] = None, backend=None, wait_for=None, marks: Optional[list] = None):
    """The stage loop of ``MVSNet.cascade`` (reference mvsnet.py:208-258) for ONE view set over the ranks of ``group`` by row
    bands.  ``ref_feats``: the reference view's feature dict (fp32); ``src_feats``: list of the source views' dicts (on CUDA:
    ``ops.HalfFeatures`` under the plain keys); every rank holds all of them.  ``wait_for(key)`` (optional) is called before a
    feature map is first used (outstanding all-gathers).  Returns per stage the re-assembled ``depth``,
    ``photometric_confidence``, ``photometric_confidence_refine``, ``depth_values_c`` and the flattened last stage like the
    reference; volumes (logits, probabilities) stay on the rank that owns the rows."""
    from . import ops
    be = backend or _CudaBackend(net)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    num_stage = net.num_stage
    rts = [ops.relative_projections(proj_matrices["stage%d" % (s + 1)]) for s in range(num_stage)]
    dev = depth_values.device
    rts = [r.to(dev, torch.float32) for r in rts]
    depth_interval = (depth_values[0, -1] - depth_values[0, 0]) / depth_values.size(1)
    outputs, last_depth = {}, None

    def mark(label):
        if marks is not None and depth_values.is_cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((label, ev))
    mark("start")
    for s in range(num_stage):
        name = "stage%d" % (s + 1)
        scale = 2 ** (num_stage - s - 1)
        h, w = int(image_hw[0]) // scale, int(image_hw[1]) // scale
        own_lo, own_hi, band_lo, band_hi = row_bands(h, world)[rank]
        owns = own_hi > own_lo
        # the sampler is a per-pixel map with a +-1 row stencil (x2 bilinear upsample): every rank evaluates it for the whole
        # image (0.1 ms at DTU size) and keeps its band
        if s == 0:
            hyp, interval = be.hypotheses_first(depth_values, net.ndepths[s], [h, w], net.inverse_depth)
        else:
            hyp, interval = be.hypotheses_next(last_depth, net.ndepths[s], net.depth_interval_ratio[s] * depth_interval, [h, w], net.inverse_depth)
        if wait_for is not None:
            wait_for(name)
        mark(name + " sampler+wait")
        d4 = hyp_c = conf = None
        if owns:
            hyp_b = hyp[:, :, band_lo:band_hi].contiguous()
            logits = be.regularised_logits(name, ref_feats[name][:, :, band_lo:band_hi], [f[name] for f in src_feats], rts[s], hyp_b, band_lo, s, False)
            d4, hyp_c, conf = be.depth_head(logits, hyp_b, interval)
            del logits
            keep = slice(own_lo - band_lo, own_hi - band_lo)
            d4, hyp_c, conf = d4[:, :, keep], hyp_c[:, :, keep], conf[:, keep]
        else:
            b = hyp.shape[0]
            d4 = hyp.new_zeros(b, 4, 0, w); hyp_c = hyp.new_zeros(b, 4, 0, w); conf = hyp.new_zeros(b, 0, w)
        mark(name + " main pass")
        packed = _assemble_rows(torch.cat([hyp_c, d4, conf.unsqueeze(1)], 1), (own_lo, own_hi), h, group)   # [B,9,h,w]
        hyp_c_full, d4_full, conf_full = packed[:, 0:4], packed[:, 4:8], packed[:, 8]
        if wait_for is not None:
            wait_for(name + "_c")
        mark(name + " exchange+wait")
        if owns:
            hyp_cb = hyp_c_full[:, :, band_lo:band_hi].contiguous()
            logits_c = be.regularised_logits(name + "_c", ref_feats[name + "_c"][:, :, band_lo:band_hi], [f[name + "_c"] for f in src_feats],
                                             rts[s], hyp_cb, band_lo, s, True)
            depth, conf_r, d4r = be.refine_head(logits_c, hyp_cb, interval)
            del logits_c
            depth, conf_r, d4r = depth[:, keep], conf_r[:, keep], d4r[:, :, keep]
        else:
            depth = hyp.new_zeros(b, 0, w); conf_r = hyp.new_zeros(b, 0, w); d4r = hyp.new_zeros(b, 4, 0, w)
        mark(name + " refine pass")
        packed = _assemble_rows(torch.cat([depth.unsqueeze(1), conf_r.unsqueeze(1), d4r], 1), (own_lo, own_hi), h, group)  # [B,6,h,w]
        mark(name + " exchange")
        stage_out = {"depth": packed[:, 0], "photometric_confidence_refine": packed[:, 1], "depth_sub_plus_refine": packed[:, 2:6],
                     "photometric_confidence": conf_full, "depth_sub_plus": d4_full, "depth_values_c": hyp_c_full,
                     "depth_values": hyp, "interval": interval}
        last_depth = stage_out["depth"]
        outputs[name] = stage_out
        outputs.update(stage_out)
    return outputs


def extract_features_view_sharded(net, imgs: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """FeatureNet for one view set ``imgs`` [1,N,3,H,W] (resident on every rank) over the ranks by VIEW: every rank computes
    the reference view and its share of the source views (round robin), then one all-gather per feature map of the fp16
    source maps.  Returns ``(ref_feats, src_feats, wait_for)``: the reference dict (fp32), the list of N-1 source dicts of
    ``ops.HalfFeatures`` and a callable that blocks the current stream on the gather a key is still waiting for."""
    from . import ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if imgs.shape[0] != 1:
        raise ValueError("single-view mode takes one view set (B = 1)")
    n = imgs.shape[1]
    mine = list(range(1 + rank, n, world))
    per = -(-(n - 1) // world)
    out = net.feature(imgs[0, [0] + mine])                       # {key: [1 + len(mine), C, h, w]} channel-last
    ref = {k: t[0:1] for k, t in out.items()}
    gathered, works = {}, {}
    for key in ("stage1", "stage1_c", "stage2", "stage2_c", "stage3", "stage3_c"):
        t = out[key]
        _, c, h, w = t.shape
        send = torch.zeros(per, h, w, c, device=t.device, dtype=torch.float16)
        if mine:
            send[:len(mine)] = ops.features_nhwc_f16(t[1:]).data
        recv = torch.empty(world * per, h, w, c, device=t.device, dtype=torch.float16)
        works[key] = dist.all_gather_into_tensor(recv, send, group=group, async_op=True)
        gathered[key] = recv
    srcs = []
    for v in range(1, n):
        j = v - 1
        slot = (j % world) * per + (j // world)
        srcs.append({k: ops.HalfFeatures(g[slot:slot + 1]) for k, g in gathered.items()})

    def wait_for(key):
        wk = works.pop(key, None)
        if wk is not None:
            wk.wait()
    return ref, srcs, wait_for


def infer_view_sharded(net, imgs: torch.Tensor, proj_matrices, depth_values: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """``MVSNet.forward`` of ONE view set on all ranks of ``group`` (every rank passes the same inputs, resident on its GPU)."""
    ref, srcs, wait_for = extract_features_view_sharded(net, imgs, group)
    return cascade_row_sharded(net, ref, srcs, proj_matrices, depth_values, imgs.shape[-2:], exchange_group(group), wait_for=wait_for)


_EXCHANGE_GROUPS = {}


def exchange_group(group: Optional[dist.ProcessGroup] = None):
    """A second communicator over the same ranks for the small per-stage exchanges: on the bulk group they would queue behind
    the six feature all-gathers (one NCCL stream per communicator, in order) - 1.7 ms of waiting at T&T size on 8 GPUs.
    Collective (every rank of ``group`` must call it the first time)."""
    key = id(group) if group is not None else None
    if key not in _EXCHANGE_GROUPS:
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
        _EXCHANGE_GROUPS[key] = dist.new_group(ranks=ranks)
    return _EXCHANGE_GROUPS[key]
