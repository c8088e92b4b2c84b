"""Golden vectors for the geometric-consistency check from the LIVE reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_fusion.py        # -> tests/golden/fusion.npz, fusion_dynamic.npz

Imports /root/reference/filter/pcd.py unmodified.  Three of its imports are unrelated to the check and absent here (plyfile,
tomlkit, yacs via filter.tank_test_config): they are stubbed.  The check itself runs torch ops after ``.cuda()``; there is no GPU in
this container, so ``Tensor.cuda`` is neutralised for the duration of the script and the SAME torch code runs on the CPU.
Inputs: depth maps of the rendered slanted plane as the rig's cameras see it (consistent by construction) + per-view noise, an
outlier patch, a zero (invalid) patch, so that every branch of the mask is exercised.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmvsnet_b200 import synthetic as syn  # noqa: E402


def fusion_case(h=48, w=64, views=4, seed=0):
    """Shared with tests/: (depths [V,h,w] fp32, intrinsics [V,3,3], extrinsics [V,4,4]) - view 0 is the reference view."""
    g = torch.Generator().manual_seed(seed)
    proj = syn.make_proj_matrices(h, w, views, 1, num_stages=3)["stage3"]
    depths = []
    for v in range(views):
        d = syn.scene_depth_view(h, w, proj, v)
        d = d * (1.0 + 0.004 * torch.randn(h, w, generator=g))      # ~0.4 % noise: straddles the 1 % relative threshold
        depths.append(d)
    depths = torch.stack(depths)
    depths[1, 5:15, 10:30] *= 1.08                                   # an inconsistent patch in one source view
    depths[2, 30:40, 40:60] = 0.0                                    # an invalid (zero) patch in another
    depths[0, 20:24, 5:12] = 0.0                                     # zeros in the reference view (pcd.py:212)
    return depths.float().contiguous(), proj[0, :, 1, :3, :3].float().contiguous(), proj[0, :, 0].float().contiguous()


def main():
    class _Any:
        def __init__(self, *a, **k): pass
        def __getattr__(self, n): return _Any()
        def __call__(self, *a, **k): return _Any()
    for name in ("plyfile", "tomlkit", "yacs", "yacs.config"):
        m = types.ModuleType(name); m.PlyData = m.PlyElement = m.value = None; m.CfgNode = _Any; sys.modules[name] = m
    m = types.ModuleType("filter.tank_test_config"); m.tank_cfg = _Any(); sys.modules["filter.tank_test_config"] = m
    sys.path.insert(0, "/root/reference")
    import filter.pcd as pcd
    torch.Tensor.cuda = lambda self, *a, **k: self                   # no GPU here: run the reference's torch code on the CPU

    depths, ks, es = fusion_case()
    views = depths.shape[0]
    ref_depth = depths[0].numpy().copy()                             # the reference patches its zeros in place, like filter_depth sees it
    out = {"depths": depths.numpy(), "intrinsics": ks.numpy(), "extrinsics": es.numpy()}
    geo_mask_sum = 0
    all_d = []
    for v in range(1, views):
        mask, d_rep, x2d, y2d = pcd.check_geometric_consistency(ref_depth, ks[0].numpy(), es[0].numpy(), depths[v].numpy().copy(),
                                                                ks[v].numpy(), es[v].numpy())
        out["mask_%d" % v], out["depth_reprojected_%d" % v], out["x2d_src_%d" % v], out["y2d_src_%d" % v] = mask, d_rep, x2d, y2d
        geo_mask_sum = geo_mask_sum + mask.astype(np.int32)          # pcd.py:291
        all_d.append(d_rep)
    out["geo_mask_sum"] = geo_mask_sum
    out["depth_est_averaged"] = (sum(all_d) + ref_depth) / (geo_mask_sum + 1)   # pcd.py:298
    path = os.path.join(ROOT, "tests", "golden", "fusion.npz")
    np.savez_compressed(path, **out)

    # dynamic-threshold variant: filter/dypcd_tanks.py check_geometric_consistency (numpy + cv2.remap) and the accumulation of
    # its filter_depth (:237-270), on the same depth maps
    import filter.dypcd_tanks as dy
    args = types.SimpleNamespace(dist_base=1 / 4, rel_diff_base=1 / 1300)
    ref_depth = depths[0].numpy().copy()
    dyn = {"dist_base": args.dist_base, "rel_diff_base": args.rel_diff_base}
    src_views = list(range(1, views))
    geo_mask_sum = 0
    dy_range = len(src_views) + 1
    geo_mask_sums = [0] * (dy_range - 2)
    all_d = []
    for v in src_views:
        masks, geo_mask, d_rep, x2d, y2d = dy.check_geometric_consistency(args, ref_depth, ks[0].numpy(), es[0].numpy(), depths[v].numpy().copy(),
                                                                          ks[v].numpy(), es[v].numpy())
        dyn["masks_%d" % v], dyn["depth_reprojected_%d" % v], dyn["x2d_src_%d" % v], dyn["y2d_src_%d" % v] = np.stack(masks), d_rep, x2d, y2d
        geo_mask_sum = geo_mask_sum + geo_mask.astype(np.int32)                  # dypcd_tanks.py:240
        for i in range(2, dy_range):
            geo_mask_sums[i - 2] = geo_mask_sums[i - 2] + masks[i - 2].astype(np.int32)
        all_d.append(d_rep)
    dyn["geo_mask_sum"] = geo_mask_sum
    dyn["depth_est_averaged"] = ((sum(all_d) + ref_depth) / (geo_mask_sum + 1)).astype(np.float32)   # :248, cast at :250
    geo_mask = geo_mask_sum >= dy_range                                          # :253
    for i in range(2, dy_range):
        geo_mask = np.logical_or(geo_mask, geo_mask_sums[i - 2] >= i)
    dyn["geo_mask"] = geo_mask
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fusion_dynamic.npz"), **dyn)
    print("dynamic: levels pass fractions", [float(dyn["masks_1"][k].mean()) for k in range(9)], "geo_mask", float(geo_mask.mean()))
    print("wrote", path, {k: (v.shape, str(v.dtype)) for k, v in out.items() if k.startswith(("mask_1", "geo", "depth_est"))},
          "consistent fraction per source:", [float(out["mask_%d" % v].mean()) for v in range(1, views)])


if __name__ == "__main__":
    main()
