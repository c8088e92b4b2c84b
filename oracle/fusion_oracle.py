"""TEST INFRASTRUCTURE ONLY - CPU restatement of the geometric-consistency check of DMVSNet's depth-map fusion.

Follows reference filter/pcd.py:152-242 (reproject_with_depth_pytorch, check_geometric_consistency[_pytorch]) and the
accumulation of filter_depth (pcd.py:283-304) with torch on the CPU, fp32, same operations in the same order.  Pinned against
the live reference by tools/make_golden_fusion.py -> tests/golden/fusion.npz (the reference module is imported there with its
unrelated dependencies stubbed and ``.cuda()`` neutralised).  Only tests/ and bench tooling may import this file.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F


def reproject_with_depth(depth_ref: torch.Tensor, k_ref: torch.Tensor, e_ref: torch.Tensor, depth_src: torch.Tensor,
                         k_src: torch.Tensor, e_src: torch.Tensor):
    """pcd.py:152-200.  Returns depth_reprojected, x_reprojected, y_reprojected [H,W] and the NORMALISED x_src, y_src [H*W]."""
    height, width = depth_ref.shape
    y_ref, x_ref = torch.meshgrid(torch.arange(0, height), torch.arange(0, width), indexing="ij")           # :160
    x_ref, y_ref = x_ref.reshape([-1]), y_ref.reshape([-1])
    xyz_ref = torch.matmul(torch.linalg.inv(k_ref), torch.vstack((x_ref, y_ref, torch.ones_like(x_ref))) * depth_ref.reshape([-1]))  # :164
    xyz_src = torch.matmul(torch.matmul(e_src, torch.linalg.inv(e_ref)), torch.vstack((xyz_ref, torch.ones_like(x_ref))))[:3]        # :167
    k_xyz_src = torch.matmul(k_src, xyz_src)                                                                   # :170
    xy_src = k_xyz_src[:2] / k_xyz_src[2:3]
    x_src = xy_src[0] / ((width - 1) / 2) - 1                                                                  # :175
    y_src = xy_src[1] / ((height - 1) / 2) - 1
    proj_xy = torch.stack((x_src, y_src), dim=-1)
    sampled = F.grid_sample(depth_src.unsqueeze(0).unsqueeze(0), proj_xy.view(1, height, width, 2), mode="bilinear",
                            padding_mode="zeros", align_corners=True).type(torch.float32).squeeze(0).squeeze(0)   # :178
    xyz_src = torch.matmul(torch.linalg.inv(k_src), torch.vstack((xy_src, torch.ones_like(x_ref))) * sampled.reshape([-1]))         # :186
    xyz_rep = torch.matmul(torch.matmul(e_ref, torch.linalg.inv(e_src)), torch.vstack((xyz_src, torch.ones_like(x_ref))))[:3]       # :189
    depth_rep = xyz_rep[2].reshape([height, width])
    k_xyz_rep = torch.matmul(k_ref, xyz_rep)
    k_xyz_rep[2:3][k_xyz_rep[2:3] == 0] += 0.00001                                                             # :194
    xy_rep = k_xyz_rep[:2] / k_xyz_rep[2:3]
    return depth_rep, xy_rep[0].reshape([height, width]), xy_rep[1].reshape([height, width]), x_src, y_src


def check_geometric_consistency(depth_ref: torch.Tensor, k_ref, e_ref, depth_src, k_src, e_src, alpha: float = 1.0):
    """pcd.py:203-224.  NOTE: like the reference it patches zeros of ``depth_ref`` IN PLACE (pcd.py:212)."""
    height, width = depth_ref.shape
    y_ref, x_ref = torch.meshgrid(torch.arange(0, height), torch.arange(0, width), indexing="ij")
    depth_rep, x_rep, y_rep, x_src, y_src = reproject_with_depth(depth_ref, k_ref, e_ref, depth_src, k_src, e_src)
    dist = torch.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)
    depth_ref[depth_ref == 0] = 1e-4
    rel = torch.abs(depth_rep - depth_ref) / depth_ref
    mask = torch.logical_and(dist < 1 * alpha, rel < 0.01 * alpha)
    depth_rep[~mask] = 0
    return mask, depth_rep, x_src, y_src


def geometric_filter(depth_ref: torch.Tensor, k_ref, e_ref, depth_srcs: Sequence[torch.Tensor], k_srcs, e_srcs,
                     thres_view: int, alpha: float = 1.0) -> Dict[str, torch.Tensor]:
    """The per-reference-view loop of filter_depth, pcd.py:283-304."""
    depth_ref = depth_ref.clone()
    geo_mask_sum = 0
    all_depth = []
    masks, xs, ys = [], [], []
    for d, k, e in zip(depth_srcs, k_srcs, e_srcs):
        mask, depth_rep, x_src, y_src = check_geometric_consistency(depth_ref, k_ref, e_ref, d, k, e, alpha)
        geo_mask_sum = geo_mask_sum + mask.to(torch.int32)
        all_depth.append(depth_rep)
        masks.append(mask); xs.append(x_src); ys.append(y_src)
    depth_avg = (sum(all_depth) + depth_ref) / (geo_mask_sum + 1)                                              # :298
    return {"geo_mask_sum": geo_mask_sum, "depth_est_averaged": depth_avg, "geo_mask": geo_mask_sum >= thres_view,
            "masks": torch.stack(masks), "depth_reprojected": torch.stack(all_depth), "x2d_src": torch.stack(xs), "y2d_src": torch.stack(ys)}


# ----------------------------------------------------------------------------- dynamic thresholds (filter/dypcd_tanks.py)
def reproject_with_depth_numpy(depth_ref, k_ref, e_ref, depth_src, k_src, e_src):
    """dypcd_tanks.py:61-98: numpy float32 inputs; int64 pixel grids promote the chain to float64; cv2.remap samples the source."""
    import cv2
    import numpy as np
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))                                       # :63
    x_ref, y_ref = x_ref.reshape([-1]), y_ref.reshape([-1])
    xyz_ref = np.matmul(np.linalg.inv(k_ref), np.vstack((x_ref, y_ref, np.ones_like(x_ref))) * depth_ref.reshape([-1]))   # :66
    xyz_src = np.matmul(np.matmul(e_src, np.linalg.inv(e_ref)), np.vstack((xyz_ref, np.ones_like(x_ref))))[:3]            # :69
    k_xyz_src = np.matmul(k_src, xyz_src)
    xy_src = k_xyz_src[:2] / k_xyz_src[2:3]                                                                     # :73
    x_src = xy_src[0].reshape([height, width]).astype(np.float32)
    y_src = xy_src[1].reshape([height, width]).astype(np.float32)
    sampled = cv2.remap(depth_src, x_src, y_src, interpolation=cv2.INTER_LINEAR)                                # :79
    xyz_src = np.matmul(np.linalg.inv(k_src), np.vstack((xy_src, np.ones_like(x_ref))) * sampled.reshape([-1]))           # :84
    xyz_rep = np.matmul(np.matmul(e_ref, np.linalg.inv(e_src)), np.vstack((xyz_src, np.ones_like(x_ref))))[:3]            # :87
    depth_rep = xyz_rep[2].reshape([height, width]).astype(np.float32)
    k_xyz_rep = np.matmul(k_ref, xyz_rep)
    k_xyz_rep[2:3][k_xyz_rep[2:3] == 0] += 0.00001
    xy_rep = k_xyz_rep[:2] / k_xyz_rep[2:3]
    return (depth_rep, xy_rep[0].reshape([height, width]).astype(np.float32), xy_rep[1].reshape([height, width]).astype(np.float32),
            x_src, y_src)


def check_geometric_consistency_dynamic(dist_base, rel_diff_base, depth_ref, k_ref, e_ref, depth_src, k_src, e_src):
    """dypcd_tanks.py:164-184: nine levels i = 2..10; returns (masks, mask, depth_reprojected, x2d_src, y2d_src)."""
    import numpy as np
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    depth_rep, x_rep, y_rep, x_src, y_src = reproject_with_depth_numpy(depth_ref, k_ref, e_ref, depth_src, k_src, e_src)
    dist = np.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(depth_rep - depth_ref) / depth_ref
        masks = [np.logical_and(dist < i * dist_base, rel < i * rel_diff_base) for i in range(2, 11)]
    mask = masks[-1]
    depth_rep[~mask] = 0
    return masks, mask, depth_rep, x_src, y_src


def geometric_filter_dynamic(dist_base, rel_diff_base, depth_ref, k_ref, e_ref, depth_srcs, k_srcs, e_srcs):
    """The per-reference-view loop of dypcd's filter_depth, dypcd_tanks.py:237-270."""
    import numpy as np
    geo_mask_sum = 0
    dy_range = len(depth_srcs) + 1
    geo_mask_sums = [0] * (dy_range - 2)
    all_depth, levels = [], []
    for d, k, e in zip(depth_srcs, k_srcs, e_srcs):
        masks, mask, depth_rep, x_src, y_src = check_geometric_consistency_dynamic(dist_base, rel_diff_base, depth_ref, k_ref, e_ref, d, k, e)
        geo_mask_sum = geo_mask_sum + mask.astype(np.int32)
        for i in range(2, dy_range):
            geo_mask_sums[i - 2] = geo_mask_sums[i - 2] + masks[i - 2].astype(np.int32)
        all_depth.append(depth_rep)
        lv = np.zeros(mask.shape, np.uint8)
        for i in range(10, 1, -1):
            lv[masks[i - 2]] = i
        levels.append(lv)
    depth_avg = (sum(all_depth) + depth_ref) / (geo_mask_sum + 1)                                              # :248
    geo_mask = geo_mask_sum >= dy_range
    for i in range(2, dy_range):
        geo_mask = np.logical_or(geo_mask, geo_mask_sums[i - 2] >= i)
    return {"geo_mask_sum": geo_mask_sum, "depth_est_averaged": depth_avg.astype(np.float32), "geo_mask": geo_mask,
            "levels": np.stack(levels), "depth_reprojected": np.stack(all_depth)}
