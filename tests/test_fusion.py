"""N4: geometric-consistency check of the depth-map fusion - oracle vs the live reference's outputs (CPU) and the
CUDA kernel vs both (GPU).  Masks are threshold decisions on fp32 chains, so a handful of pixels within an ulp of a threshold may
flip between two fp32 implementations: the tests bound the disagreement instead of demanding equality."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_fusion import fusion_case  # noqa: E402
from oracle import fusion_oracle as FO  # noqa: E402


def _golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "fusion.npz"))
    return {k: z[k] for k in z.files}


def test_fixture_inputs_regenerate():
    g = _golden()
    depths, ks, es = fusion_case()
    assert np.array_equal(depths.numpy(), g["depths"]) and np.array_equal(ks.numpy(), g["intrinsics"]) and np.array_equal(es.numpy(), g["extrinsics"])


def test_oracle_matches_live_reference_outputs():
    g = _golden()
    depths, ks, es = fusion_case()
    views = depths.shape[0]
    out = FO.geometric_filter(depths[0], ks[0], es[0], [depths[v] for v in range(1, views)], [ks[v] for v in range(1, views)],
                              [es[v] for v in range(1, views)], thres_view=2)
    for i, v in enumerate(range(1, views)):
        want_mask = torch.from_numpy(g["mask_%d" % v])
        agree = float((out["masks"][i] == want_mask).float().mean())
        assert agree >= 0.999, (v, agree)
        both = out["masks"][i] & want_mask
        got_d, want_d = out["depth_reprojected"][i][both], torch.from_numpy(g["depth_reprojected_%d" % v])[both]
        assert float(((got_d - want_d).abs() / want_d.abs()).max()) < 1e-5
        want_x = torch.from_numpy(g["x2d_src_%d" % v])
        assert float(((out["x2d_src"][i] - want_x).abs() / want_x.abs().clamp_min(1.0)).max()) < 1e-4
    assert float((out["geo_mask_sum"] != torch.from_numpy(g["geo_mask_sum"])).float().mean()) <= 0.002
    same = out["geo_mask_sum"] == torch.from_numpy(g["geo_mask_sum"])
    rel = ((out["depth_est_averaged"].double() - torch.from_numpy(g["depth_est_averaged"])).abs() / torch.from_numpy(g["depth_est_averaged"]).abs())[same]
    assert float(rel.max()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,views", [(48, 64, 4), (96, 160, 6)])
def test_geo_consistency_kernel_vs_oracle(h, w, views):
    from dmvsnet_b200 import fusion
    depths, ks, es = fusion_case(h, w, views, seed=1)
    srcs = list(range(1, views))
    want = FO.geometric_filter(depths[0], ks[0], es[0], [depths[v] for v in srcs], [ks[v] for v in srcs], [es[v] for v in srcs], thres_view=2)
    got = fusion.geometric_filter(depths[0].cuda(), ks[0], es[0], [depths[v].cuda() for v in srcs], [ks[v] for v in srcs],
                                  [es[v] for v in srcs], thres_view=2, per_source=True)
    masks = got["masks"].cpu()
    assert float((masks == want["masks"]).float().mean()) >= 0.998
    both = masks & want["masks"]
    assert 0.3 < float(both.float().mean()) < 0.99                      # the fixture has consistent and inconsistent regions
    rel = ((got["depth_reprojected"].cpu() - want["depth_reprojected"]).abs() / want["depth_reprojected"].abs().clamp_min(1e-6))[both]
    assert float(rel.max()) < 1e-5
    assert float((got["depth_reprojected"].cpu()[~masks]).abs().max()) == 0.0
    assert float(((got["x2d_src"].cpu() - want["x2d_src"]).abs() / want["x2d_src"].abs().clamp_min(1.0)).max()) < 1e-4
    same = got["geo_mask_sum"].cpu() == want["geo_mask_sum"]
    assert float(same.float().mean()) >= 0.995
    rel = ((got["depth_est_averaged"].cpu() - want["depth_est_averaged"]).abs() / want["depth_est_averaged"].abs())[same]
    assert float(rel.max()) < 1e-5
    assert torch.equal(got["geo_mask"].cpu()[same], want["geo_mask"][same])
    # the reference-signature entry points (one source view)
    m, d, x, y = fusion.check_geometric_consistency(depths[0].numpy(), ks[0].numpy(), es[0].numpy(), depths[1].numpy(), ks[1].numpy(), es[1].numpy())
    assert isinstance(m, np.ndarray) and m.dtype == np.bool_ and d.shape == (h, w) and x.shape == (h * w,)
    assert np.array_equal(m, masks[0].numpy()) and np.array_equal(d, got["depth_reprojected"][0].cpu().numpy())


@pytest.mark.gpu
def test_geo_consistency_kernel_vs_live_reference_fixture():
    from dmvsnet_b200 import fusion
    g = _golden()
    depths, ks, es = fusion_case()
    srcs = [1, 2, 3]
    got = fusion.geometric_filter(depths[0].cuda(), ks[0], es[0], [depths[v].cuda() for v in srcs], [ks[v] for v in srcs],
                                  [es[v] for v in srcs], thres_view=2, per_source=True)
    for i, v in enumerate(srcs):
        assert float((got["masks"][i].cpu() == torch.from_numpy(g["mask_%d" % v])).float().mean()) >= 0.998
    same = got["geo_mask_sum"].cpu() == torch.from_numpy(g["geo_mask_sum"])
    assert float(same.float().mean()) >= 0.995
    rel = ((got["depth_est_averaged"].cpu().double() - torch.from_numpy(g["depth_est_averaged"])).abs() / torch.from_numpy(g["depth_est_averaged"]).abs())[same]
    assert float(rel.max()) < 1e-5


# ----------------------------------------------------------------------------- dynamic thresholds (filter/dypcd_tanks.py)
def _golden_dynamic():
    z = np.load(os.path.join(ROOT, "tests", "golden", "fusion_dynamic.npz"))
    return {k: z[k] for k in z.files}


def _np_case(h=48, w=64, views=4, seed=0):
    depths, ks, es = fusion_case(h, w, views, seed=seed)
    return depths.numpy(), ks.numpy(), es.numpy()


def test_dynamic_oracle_matches_live_reference_outputs():
    """Same numpy / cv2 calls in the same order as dypcd_tanks.py: the restatement reproduces the live reference's outputs exactly."""
    g = _golden_dynamic()
    depths, ks, es = _np_case()
    srcs = [1, 2, 3]
    out = FO.geometric_filter_dynamic(float(g["dist_base"]), float(g["rel_diff_base"]), depths[0].copy(), ks[0], es[0],
                                      [depths[v].copy() for v in srcs], [ks[v] for v in srcs], [es[v] for v in srcs])
    for i, v in enumerate(srcs):
        for k in range(9):
            lv = out["levels"][i]
            assert np.array_equal(np.logical_and(lv != 0, lv <= k + 2), g["masks_%d" % v][k])
        assert np.array_equal(out["depth_reprojected"][i], g["depth_reprojected_%d" % v])
    assert np.array_equal(out["geo_mask_sum"], g["geo_mask_sum"]) and np.array_equal(out["geo_mask"], g["geo_mask"])
    assert np.array_equal(out["depth_est_averaged"], g["depth_est_averaged"])


def _check_dynamic(got, want_levels, want_drep, want_sum, want_mask, want_avg):
    """Threshold decisions on (float64 / float32) chains evaluated in a different operation order: allow a 0.2 % disagreement on
    the levels, compare values where the decisions agree."""
    lv = got["levels"].cpu().numpy()
    agree = lv == want_levels
    assert agree.mean() >= 0.998, agree.mean()
    drep = got["depth_reprojected"].cpu().numpy()
    both = np.logical_and(agree, want_levels != 0)
    assert 0.2 < both.mean() < 0.99
    assert np.abs(drep[both] - want_drep[both]).max() <= 1e-6 * np.abs(want_drep[both]).max()
    assert np.abs(drep[lv == 0]).max() == 0.0
    same = got["geo_mask_sum"].cpu().numpy() == want_sum
    assert same.mean() >= 0.995
    pix_ok = agree.all(0)                                    # pixels whose every per-source decision agrees
    assert np.array_equal(got["geo_mask"].cpu().numpy()[pix_ok], want_mask[pix_ok])
    avg = got["depth_est_averaged"].cpu().numpy()
    assert np.abs(avg[pix_ok] - want_avg[pix_ok]).max() <= 1e-6 * np.abs(want_avg[pix_ok]).max()


@pytest.mark.gpu
def test_dynamic_kernel_vs_live_reference_fixture():
    from dmvsnet_b200 import fusion
    g = _golden_dynamic()
    depths, ks, es = _np_case()
    srcs = [1, 2, 3]
    got = fusion.geometric_filter_dynamic(depths[0], ks[0], es[0], [depths[v] for v in srcs], [ks[v] for v in srcs], [es[v] for v in srcs],
                                          float(g["dist_base"]), float(g["rel_diff_base"]), per_source=True)
    want_levels = np.zeros((3,) + depths[0].shape, np.uint8)
    for i, v in enumerate(srcs):
        for k in range(8, -1, -1):
            want_levels[i][g["masks_%d" % v][k]] = k + 2
    _check_dynamic(got, want_levels, np.stack([g["depth_reprojected_%d" % v] for v in srcs]), g["geo_mask_sum"], g["geo_mask"], g["depth_est_averaged"])
    for i, v in enumerate(srcs):                              # the float32 pixel coordinates handed to cv2.remap
        assert np.abs(got["x2d_src"][i].cpu().numpy() - g["x2d_src_%d" % v]).max() <= 1e-4
        assert np.abs(got["y2d_src"][i].cpu().numpy() - g["y2d_src_%d" % v]).max() <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,views,dist_base,rel_base", [(48, 64, 4, 0.25, 1 / 1300), (96, 160, 6, 0.5, 1 / 800), (40, 56, 11, 0.25, 1 / 1300)])
def test_dynamic_kernel_vs_oracle(h, w, views, dist_base, rel_base):
    import types
    from dmvsnet_b200 import fusion
    depths, ks, es = _np_case(h, w, views, seed=2)
    srcs = list(range(1, views))
    want = FO.geometric_filter_dynamic(dist_base, rel_base, depths[0].copy(), ks[0], es[0], [depths[v].copy() for v in srcs],
                                       [ks[v] for v in srcs], [es[v] for v in srcs])
    got = fusion.geometric_filter_dynamic(torch.from_numpy(depths[0]).cuda(), ks[0], es[0], [depths[v] for v in srcs], [ks[v] for v in srcs],
                                          [es[v] for v in srcs], dist_base, rel_base, per_source=True)
    _check_dynamic(got, want["levels"], want["depth_reprojected"], want["geo_mask_sum"], want["geo_mask"], want["depth_est_averaged"])
    # the reference-signature entry point (one source view, numpy results)
    args = types.SimpleNamespace(dist_base=dist_base, rel_diff_base=rel_base)
    masks, mask, drep, x2d, y2d = fusion.check_geometric_consistency_dynamic(args, depths[0], ks[0], es[0], depths[1], ks[1], es[1])
    assert len(masks) == 9 and masks[0].dtype == np.bool_ and mask is masks[-1] and drep.shape == (h, w) and x2d.shape == (h, w)
    lv = got["levels"][0].cpu().numpy()
    assert all(np.array_equal(masks[k], np.logical_and(lv != 0, lv <= k + 2)) for k in range(9))
    assert all((masks[k] <= masks[k + 1]).all() for k in range(8))          # the levels are nested


@pytest.mark.gpu
def test_dynamic_fused_mask_rejects_more_than_ten_sources():
    from dmvsnet_b200 import fusion
    depths, ks, es = _np_case(16, 24, 12, seed=0)
    srcs = list(range(1, 12))
    with pytest.raises(Exception, match="at most 10 source views"):
        fusion.geometric_filter_dynamic(depths[0], ks[0], es[0], [depths[v] for v in srcs], [ks[v] for v in srcs], [es[v] for v in srcs], 0.25, 1 / 1300)


# ----------------------------------------------------------------------------- filter_depth on a scan folder
def _dense(points, colors, masks):
    """Scatter the concatenated per-view point lists back onto their pixels: [V,H,W,3] maps (row-major selection order)."""
    v, h, w = masks.shape
    xyz = np.zeros((v, h, w, 3), np.float32)
    rgb = np.zeros((v, h, w, 3), np.uint8)
    at = 0
    for i in range(v):
        n = int(masks[i].sum())
        xyz[i][masks[i]] = points[at:at + n]
        rgb[i][masks[i]] = colors[at:at + n]
        at += n
    assert at == len(points)
    return xyz, rgb


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["static", "dynamic"])
def test_filter_depth_scan_folder_vs_live_reference(tag, tmp_path):
    """The whole fusion step on the committed scan folder against what the live reference's filter_depth produced from it
    (tools/make_golden_scene.py): masks, vertices, colours, PLY file."""
    import shutil
    import types
    from PIL import Image
    from dmvsnet_b200 import formats, fusion
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_scene import ARGS
    want = np.load(os.path.join(ROOT, "tests", "golden", "fusion_scene_expected.npz"))
    work = str(tmp_path / "scan")
    shutil.copytree(os.path.join(ROOT, "tests", "golden", "fusion_scene"), work)
    ply = str(tmp_path / "pcd" / "out.ply")
    points, colors = fusion.filter_depth(types.SimpleNamespace(**ARGS), work, work, work, ply, dynamic=(tag == "dynamic"))
    got_masks = {k: np.stack([np.array(Image.open(os.path.join(work, "mask/{:0>8}_{}.png".format(v, k)))) > 0 for v in range(4)])
                 for k in ("photo", "geo", "final")}
    want_masks = {k: np.stack([want["%s_mask_%s_%d" % (tag, k, v)] for v in range(4)]) for k in ("photo", "geo", "final")}
    assert np.array_equal(got_masks["photo"], want_masks["photo"])
    for k in ("geo", "final"):
        assert (got_masks[k] == want_masks[k]).mean() >= 0.995, (k, (got_masks[k] == want_masks[k]).mean())
    got_xyz, got_rgb = _dense(points, colors, got_masks["final"])
    want_xyz, want_rgb = _dense(want[tag + "_xyz"], want[tag + "_rgb"], want_masks["final"])
    both = got_masks["final"] & want_masks["final"]
    assert both.sum() > 300
    scale = np.abs(want_xyz[both]).max()
    assert np.abs(got_xyz[both] - want_xyz[both]).max() <= 2e-6 * scale
    assert np.array_equal(got_rgb[both], want_rgb[both])
    p2, c2 = formats.read_ply(ply)                            # what was written is what was returned
    assert np.array_equal(p2, points) and np.array_equal(c2, colors)
    if tag == "dynamic":
        assert os.path.exists(os.path.join(work, "depth_est/00000000_averaged.pfm"))
