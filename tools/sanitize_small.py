"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): native FeatureNet + 3-stage cascade from images,
all three W1 kernels, B = 2; the fusion kernels and the W1 backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmvsnet_b200 import MVSNet, ops, synthetic as syn

b, n, h, w, nd, ratios = 2, 3, 64, 96, [16, 8, 8], [4, 2, 1]
net = MVSNet(nd, ratios, inverse_depth=True)
net.load_state_dict(syn.randomise_regnet_state(net.state_dict(), seed=5))
net = net.cuda().eval()
imgs = syn.make_images(h, w, n, b, seed=6)
proj = syn.make_proj_matrices(h, w, n, b, num_stages=3)
dv = syn.make_depth_values(b, 192, inverse=True)
with torch.no_grad():
    out = net(imgs.cuda(), proj, dv.cuda())
    torch.cuda.synchronize()
    host = net.infer(imgs[:1], {k: v[:1] for k, v in proj.items()}, dv[:1])
    g = torch.Generator().manual_seed(1)
    feats = [torch.randn(1, 16, 40, 72, generator=g).cuda() for _ in range(4)]
    rt = ops.relative_projections(syn.make_proj_matrices(160, 288, 4, 1, num_stages=1)["stage1"]).cuda()
    hyp = (425 + 500 * torch.rand(1, 8, 40, 72, generator=g)).cuda()
    ref = None
    for layout in ("nchw", "nhwc", "staged"):
        c = ops.warp_corr(feats, rt, hyp, layout=layout)
        ref = c if ref is None else ref
        assert float((c - ref).abs().max()) < 1e-4
    torch.cuda.synchronize()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from make_golden_fusion import fusion_case
    from dmvsnet_b200 import fusion
    depths, ks, es = fusion_case(40, 56, 4, seed=3)
    fz = fusion.geometric_filter(depths[0].cuda(), ks[0], es[0], [depths[v].cuda() for v in (1, 2, 3)], [ks[v] for v in (1, 2, 3)],
                                 [es[v] for v in (1, 2, 3)], thres_view=2, per_source=True)
    fd = fusion.geometric_filter_dynamic(depths[0].cuda(), ks[0], es[0], [depths[v] for v in (1, 2, 3)], [ks[v] for v in (1, 2, 3)],
                                         [es[v] for v in (1, 2, 3)], 0.25, 1 / 1300, per_source=True)
    xyz = fusion.backproject_world(fd["depth_est_averaged"], ks[0], es[0])
    # W1 backward (odd sizes, every channel count)
    for cch in (8, 16, 32):
        fb = [torch.randn(1, cch, 21, 37, generator=g).cuda() for _ in range(3)]
        rtb = ops.relative_projections(syn.make_proj_matrices(84, 148, 3, 1, num_stages=1)["stage1"]).cuda()
        hb = (425 + 500 * torch.rand(1, 5, 21, 37, generator=g)).cuda()
        gb = ops.warp_corr_backward(fb, rtb, hb, torch.randn(1, 2, 5, 21, 37, generator=g).cuda())
        assert all(bool(torch.isfinite(t).all()) for t in gb)
    torch.cuda.synchronize()
print("ok", int(fd["geo_mask"].sum()), float(xyz.abs().max()), float(out["depth"].mean()), float(host["depth"].mean()), int(fz["geo_mask_sum"].sum()))
