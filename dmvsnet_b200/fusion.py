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


@ops._on_device
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


@ops._on_device
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


# ----------------------------------------------------------------------------- point cloud (filter_depth)
@torch.no_grad()
@ops._on_device
def backproject_world(depth, intrinsics, extrinsics) -> torch.Tensor:
    """depth [H,W] -> world points [H,W,3] float32 for every pixel (pcd.py:340-343; float64 chain inside the kernel)."""
    lib = N.load()
    dev = depth.device if isinstance(depth, torch.Tensor) and depth.is_cuda else torch.device("cuda", torch.cuda.current_device())
    d = _t(depth, dev).contiguous()
    k, e = _np32(intrinsics), _np32(extrinsics)
    mats = torch.from_numpy(np.concatenate([np.linalg.inv(k[:3, :3]).reshape(-1), np.linalg.inv(e)[:3, :4].reshape(-1)]).astype(np.float32)).to(dev)
    h, w = d.shape
    xyz = torch.empty(h, w, 3, device=dev, dtype=torch.float32)
    N.check(lib.dmvs_backproject_world_f32(d.data_ptr(), mats.data_ptr(), h, w, xyz.data_ptr(), ops._stream()), "dmvs_backproject_world_f32")
    return xyz


def _read_img(filename: str) -> np.ndarray:
    from PIL import Image
    return np.array(Image.open(filename), dtype=np.float32) / 255.0          # pcd.py:44-48


def _save_mask(filename: str, mask: np.ndarray) -> None:
    from PIL import Image
    Image.fromarray(mask.astype(np.uint8) * 255).save(filename)                # pcd.py:36-40


def filter_depth(args, pair_folder: str, scan_folder: str, out_folder: str, plyfilename: str, dynamic: bool = False, verbose: bool = False):
    """The reference's ``filter_depth`` (filter/pcd.py:244-361; with ``dynamic`` the dypcd form, filter/dypcd_tanks.py:186-326) on the
    GPU: same folder layout in (``pair.txt``, ``cams/*_cam.txt``, ``images/*.jpg``, ``depth_est/*.pfm``, ``confidence/*[_stageK].pfm``)
    and out (``mask/*_{photo,geo,final}.png``, the PLY; ``dynamic`` also writes ``depth_est/*_averaged.pfm``), same ``args`` fields
    (``ndepths``, ``conf``, ``thres_view`` | ``dist_base``, ``rel_diff_base``).  Per reference view: ONE launch for the geometric check
    over all its source views, one for the back-projection; masks, selection and colours are torch ops on the device.
    Returns (points [N,3] float32, colors [N,3] uint8) as numpy arrays, in the order they are written."""
    import os

    from . import formats

    num_stage = len(args.ndepths)
    dev = torch.device("cuda", torch.cuda.current_device())
    pts, cols = [], []
    cam = lambda v: formats.read_camera_parameters(os.path.join(scan_folder, "cams/{:0>8}_cam.txt".format(v)))  # noqa: E731
    pfm = lambda sub, name: np.ascontiguousarray(formats.read_pfm(os.path.join(out_folder, sub, name))[0])  # noqa: E731
    for ref_view, src_views in formats.read_pair_file(os.path.join(pair_folder, "pair.txt")):
        k_ref, e_ref = cam(ref_view)
        ref_img = _read_img(os.path.join(scan_folder, "images/{:0>8}.jpg".format(ref_view)))
        ref_depth = pfm("depth_est", "{:0>8}.pfm".format(ref_view))
        conf = pfm("confidence", "{:0>8}.pfm".format(ref_view))
        if os.path.exists(os.path.join(out_folder, "confidence/{:0>8}_stage2.pfm".format(ref_view))):
            conf2 = pfm("confidence", "{:0>8}_stage2.pfm".format(ref_view))
            conf1 = pfm("confidence", "{:0>8}_stage1.pfm".format(ref_view))
        else:
            conf2 = conf1 = conf
        c, c2, c1 = [torch.from_numpy(a).to(dev) for a in (conf, conf2, conf1)]
        photo_mask = (c > args.conf[2]) & (c2 > args.conf[1]) & (c1 > args.conf[0])                 # pcd.py:273
        cams = [cam(v) for v in src_views]
        depths = [pfm("depth_est", "{:0>8}.pfm".format(v)) for v in src_views]
        if dynamic:
            out = geometric_filter_dynamic(ref_depth, k_ref, e_ref, depths, [k for k, _ in cams], [e for _, e in cams], args.dist_base, args.rel_diff_base)
            formats.save_pfm(os.path.join(out_folder, "depth_est/{:0>8}_averaged.pfm".format(ref_view)), out["depth_est_averaged"].cpu().numpy())
        else:
            out = geometric_filter(ref_depth, k_ref, e_ref, depths, [k for k, _ in cams], [e for _, e in cams], args.thres_view)
        geo_mask = out["geo_mask"]
        final_mask = photo_mask & geo_mask
        os.makedirs(os.path.join(out_folder, "mask"), exist_ok=True)
        for tag, m in (("photo", photo_mask), ("geo", geo_mask), ("final", final_mask)):
            _save_mask(os.path.join(out_folder, "mask/{:0>8}_{}.png".format(ref_view, tag)), m.cpu().numpy())
        if verbose:
            print("processing {}, ref-view{:0>2}, photo/geo/final-mask:{}/{}/{}".format(scan_folder, ref_view, float(photo_mask.float().mean()),
                                                                                        float(geo_mask.float().mean()), float(final_mask.float().mean())))
        if num_stage == 1:
            img = ref_img[1::4, 1::4, :]
        elif num_stage == 2:
            img = ref_img[1::2, 1::2, :]
        else:
            img = ref_img
        xyz = backproject_world(out["depth_est_averaged"], k_ref, e_ref)
        pts.append(xyz[final_mask].cpu().numpy())
        color = torch.from_numpy(np.ascontiguousarray(img)).to(dev)[final_mask]
        cols.append((color * 255).to(torch.uint8).cpu().numpy())                                    # pcd.py:345: truncation, not rounding
    points, colors = np.concatenate(pts, 0), np.concatenate(cols, 0)
    os.makedirs(os.path.dirname(os.path.abspath(plyfilename)), exist_ok=True)
    formats.write_ply(plyfilename, points, colors)
    if verbose:
        print("saving the final model to", plyfilename)
    return points, colors
