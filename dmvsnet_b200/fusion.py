"""Geometric-consistency check of the depth-map fusion on the GPU - the step that consumes the path's depth
maps.  Mirrors reference filter/pcd.py: ``reproject_with_depth_pytorch`` / ``check_geometric_consistency_pytorch`` (pcd.py:152-224,
same argument order, torch tensors back), ``check_geometric_consistency`` (pcd.py:226-242, numpy back) and, as one fused launch
over all source views, the accumulation loop of ``filter_depth`` (pcd.py:283-304) -> ``geometric_filter``.

Everything per pixel runs in ``dmvs_geo_consistency_f32``; the six 3x3 / 3x4 matrices per source view are computed here with
the reference's own torch calls (``torch.linalg.inv`` / ``matmul`` in fp32 on the CPU), so they are bit-equal to the
reference's on the same host.  No CPU fallback: the depth maps must be (or are moved to) CUDA tensors.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from . import ops


def _t(x, dev=None) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    t = t.to(torch.float32)
    return t.to(dev) if dev is not None else t


def source_matrices(k_ref, e_ref, k_src, e_src) -> torch.Tensor:
    """[60] floats for one source view: inv(K_ref) | (E_src inv(E_ref))[:3,:4] | K_src | inv(K_src) | (E_ref inv(E_src))[:3,:4] | K_ref,
    the matrices of pcd.py:164-191, computed like the reference does (fp32, torch CPU)."""
    k_ref, e_ref, k_src, e_src = [_t(m).cpu() for m in (k_ref, e_ref, k_src, e_src)]
    t1 = torch.matmul(e_src, torch.linalg.inv(e_ref))
    t2 = torch.matmul(e_ref, torch.linalg.inv(e_src))
    parts = [torch.linalg.inv(k_ref[:3, :3]), t1[:3, :4], k_src[:3, :3], torch.linalg.inv(k_src[:3, :3]), t2[:3, :4], k_ref[:3, :3]]
    return torch.cat([p.reshape(-1) for p in parts]).contiguous()


def _run(depth_ref, mats, depth_srcs, alpha, want_per_source, want_fused):
    lib = N.load()
    dev = depth_srcs.device
    s, h, w = depth_srcs.shape
    mask = torch.empty(s, h, w, device=dev, dtype=torch.uint8) if want_per_source else None
    drep = torch.empty(s, h, w, device=dev, dtype=torch.float32) if want_per_source else None
    xy = torch.empty(s, 2, h, w, device=dev, dtype=torch.float32) if want_per_source else None
    msum = torch.empty(h, w, device=dev, dtype=torch.int32) if want_fused else None
    davg = torch.empty(h, w, device=dev, dtype=torch.float32) if want_fused else None
    rc = lib.dmvs_geo_consistency_f32(depth_ref.data_ptr(), depth_srcs.data_ptr(), mats.data_ptr(), s, h, w, 1.0 * alpha, 0.01 * alpha,
                                      ops._ptr(mask), ops._ptr(drep), ops._ptr(xy), ops._ptr(msum), ops._ptr(davg), ops._stream())
    N.check(rc, "dmvs_geo_consistency_f32")
    return mask, drep, xy, msum, davg


def _prepare(depth_ref, k_ref, e_ref, depth_srcs, k_srcs, e_srcs):
    dev = depth_ref.device if isinstance(depth_ref, torch.Tensor) and depth_ref.is_cuda else torch.device("cuda", torch.cuda.current_device())
    dref = _t(depth_ref, dev).contiguous()
    dsrc = torch.stack([_t(d, dev) for d in depth_srcs]).contiguous()
    if dsrc.shape[1:] != dref.shape:
        raise ValueError("all depth maps must share one resolution, got %s vs %s" % (tuple(dsrc.shape[1:]), tuple(dref.shape)))
    mats = torch.stack([source_matrices(k_ref, e_ref, k, e) for k, e in zip(k_srcs, e_srcs)]).to(dev)
    return dref, dsrc, mats


@torch.no_grad()
def check_geometric_consistency_pytorch(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src, alpha=1.0):
    """pcd.py:203-224: (mask [H,W] bool, depth_reprojected [H,W] with 0 where inconsistent, x2d_src, y2d_src [H*W] normalised) as CUDA
    tensors.  Unlike the reference it does not patch zeros of the caller's ``depth_ref`` in place (the 1e-4 substitution of
    pcd.py:212 is applied inside the kernel)."""
    dref, dsrc, mats = _prepare(depth_ref, intrinsics_ref, extrinsics_ref, [depth_src], [intrinsics_src], [extrinsics_src])
    mask, drep, xy, _, _ = _run(dref, mats, dsrc, alpha, True, False)
    return mask[0].bool(), drep[0], xy[0, 0].reshape(-1), xy[0, 1].reshape(-1)


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    """pcd.py:226-242: the numpy-returning form ``filter_depth`` calls."""
    mask, drep, x, y = check_geometric_consistency_pytorch(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src)
    return mask.cpu().numpy(), drep.cpu().numpy(), x.cpu().numpy(), y.cpu().numpy()


@torch.no_grad()
def geometric_filter(depth_ref, intrinsics_ref, extrinsics_ref, depth_srcs: Sequence, intrinsics_srcs: Sequence, extrinsics_srcs: Sequence,
                     thres_view: int, alpha: float = 1.0, per_source: bool = False) -> Dict[str, torch.Tensor]:
    """The source-view loop of ``filter_depth`` (pcd.py:283-304) as ONE kernel launch: ``geo_mask_sum`` [H,W] int32,
    ``depth_est_averaged`` [H,W], ``geo_mask`` = geo_mask_sum >= thres_view; with ``per_source`` also the per-view masks / reprojected
    depths / normalised source coordinates."""
    dref, dsrc, mats = _prepare(depth_ref, intrinsics_ref, extrinsics_ref, depth_srcs, intrinsics_srcs, extrinsics_srcs)
    mask, drep, xy, msum, davg = _run(dref, mats, dsrc, alpha, per_source, True)
    out = {"geo_mask_sum": msum, "depth_est_averaged": davg, "geo_mask": msum >= int(thres_view)}
    if per_source:
        out.update({"masks": mask.bool(), "depth_reprojected": drep, "x2d_src": xy[:, 0].reshape(len(dsrc), -1), "y2d_src": xy[:, 1].reshape(len(dsrc), -1)})
    return out


# ----------------------------------------------------------------------------- dynamic thresholds (filter/dypcd_tanks.py)
def _np32(x) -> np.ndarray:
    return np.asarray(x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x, dtype=np.float32)


def source_matrices_numpy(k_ref, e_ref, k_src, e_src) -> torch.Tensor:
    """The same six matrices as ``source_matrices``, computed like filter/dypcd_tanks.py:66-91 does: numpy float32
    ``np.linalg.inv`` / ``np.matmul`` (the kernel promotes them to float64 exactly, as numpy does when it multiplies them into
    float64 point arrays)."""
    k_ref, e_ref, k_src, e_src = [_np32(m) for m in (k_ref, e_ref, k_src, e_src)]
    t1 = np.matmul(e_src, np.linalg.inv(e_ref))
    t2 = np.matmul(e_ref, np.linalg.inv(e_src))
    parts = [np.linalg.inv(k_ref[:3, :3]), t1[:3, :4], k_src[:3, :3], np.linalg.inv(k_src[:3, :3]), t2[:3, :4], k_ref[:3, :3]]
    return torch.from_numpy(np.concatenate([np.asarray(p, np.float32).reshape(-1) for p in parts]))


def _run_dynamic(depth_ref, mats, depth_srcs, dist_base, rel_diff_base, want_per_source, want_fused):
    lib = N.load()
    dev = depth_srcs.device
    s, h, w = depth_srcs.shape
    level = torch.empty(s, h, w, device=dev, dtype=torch.uint8) if want_per_source else None
    drep = torch.empty(s, h, w, device=dev, dtype=torch.float32) if want_per_source else None
    xy = torch.empty(s, 2, h, w, device=dev, dtype=torch.float32) if want_per_source else None
    msum = torch.empty(h, w, device=dev, dtype=torch.int32) if want_fused else None
    gmask = torch.empty(h, w, device=dev, dtype=torch.uint8) if want_fused else None
    davg = torch.empty(h, w, device=dev, dtype=torch.float32) if want_fused else None
    rc = lib.dmvs_geo_consistency_dynamic_f32(depth_ref.data_ptr(), depth_srcs.data_ptr(), mats.data_ptr(), s, h, w, float(dist_base),
                                              float(rel_diff_base), ops._ptr(level), ops._ptr(drep), ops._ptr(xy), ops._ptr(msum),
                                              ops._ptr(gmask), ops._ptr(davg), ops._stream())
    N.check(rc, "dmvs_geo_consistency_dynamic_f32")
    return level, drep, xy, msum, gmask, davg


def _prepare_dynamic(depth_ref, k_ref, e_ref, depth_srcs, k_srcs, e_srcs):
    dev = depth_ref.device if isinstance(depth_ref, torch.Tensor) and depth_ref.is_cuda else torch.device("cuda", torch.cuda.current_device())
    dref = _t(depth_ref, dev).contiguous()
    dsrc = torch.stack([_t(d, dev) for d in depth_srcs]).contiguous()
    if dsrc.shape[1:] != dref.shape:
        raise ValueError("all depth maps must share one resolution, got %s vs %s" % (tuple(dsrc.shape[1:]), tuple(dref.shape)))
    mats = torch.stack([source_matrices_numpy(k_ref, e_ref, k, e) for k, e in zip(k_srcs, e_srcs)]).to(dev)
    return dref, dsrc, mats


@torch.no_grad()
def check_geometric_consistency_dynamic(args, depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    """filter/dypcd_tanks.py:164-184, same argument order (``args`` carries ``dist_base`` and ``rel_diff_base``) and numpy
    results: (masks - nine bool maps for the levels i = 2..10 -, mask = masks[-1], depth_reprojected, x2d_src, y2d_src [H,W]
    float32 pixel coordinates)."""
    dref, dsrc, mats = _prepare_dynamic(depth_ref, intrinsics_ref, extrinsics_ref, [depth_src], [intrinsics_src], [extrinsics_src])
    level, drep, xy, _, _, _ = _run_dynamic(dref, mats, dsrc, args.dist_base, args.rel_diff_base, True, False)
    lv = level[0].cpu().numpy()
    masks = [np.logical_and(lv != 0, lv <= i) for i in range(2, 11)]
    return masks, masks[-1], drep[0].cpu().numpy(), xy[0, 0].cpu().numpy(), xy[0, 1].cpu().numpy()


@torch.no_grad()
def geometric_filter_dynamic(depth_ref, intrinsics_ref, extrinsics_ref, depth_srcs: Sequence, intrinsics_srcs: Sequence,
                             extrinsics_srcs: Sequence, dist_base: float, rel_diff_base: float, per_source: bool = False) -> Dict[str, torch.Tensor]:
    """The source-view loop of dypcd's ``filter_depth`` (dypcd_tanks.py:237-270) as ONE kernel launch: ``geo_mask_sum`` (sources
    passing the loosest level), ``depth_est_averaged`` and ``geo_mask`` = OR over i = 2..S of (#sources passing level i) >= i; with
    ``per_source`` also ``levels`` [S,H,W] uint8 (smallest level passed, 0 = none), the reprojected depths and the source pixel
    coordinates."""
    dref, dsrc, mats = _prepare_dynamic(depth_ref, intrinsics_ref, extrinsics_ref, depth_srcs, intrinsics_srcs, extrinsics_srcs)
    level, drep, xy, msum, gmask, davg = _run_dynamic(dref, mats, dsrc, dist_base, rel_diff_base, per_source, True)
    out = {"geo_mask_sum": msum, "depth_est_averaged": davg, "geo_mask": gmask.bool()}
    if per_source:
        out.update({"levels": level, "depth_reprojected": drep, "x2d_src": xy[:, 0], "y2d_src": xy[:, 1]})
    return out
