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
