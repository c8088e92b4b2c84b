"""Parity of the CUDA path (through the C ABI) against the oracle and the committed reference outputs.

Tolerances (fp32 path): cost volume 2e-5 of max|cost| (sum re-association only); single conv layers 1e-5;
logits 1e-4 of max|logit|; regressed depths / final depth 1e-3 relative (BASELINE.json north_star), measured
far tighter in practice - the assert message prints the achieved error.
"""
import pytest
import torch

from cases import CASES, case_inputs, case_state
from conftest import load_golden, rel_linf
from oracle import dmvs_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _lib(native_lib):
    return native_lib


def cuda(t):
    return t.to(DEV)


# ------------------------------------------------------------------------------------------ W1
@pytest.fixture(params=["nhwc", "nchw", "staged", "auto"])
def w1_layout(request):
    """All W1 kernels (channel-last gather, reference layout, TMA-staged with fallback pass) go through every W1 test."""
    from dmvsnet_b200 import ops
    old, ops.W1_LAYOUT = ops.W1_LAYOUT, request.param
    yield request.param
    ops.W1_LAYOUT = old


def test_warp_corr_edge_fixture(w1_layout):
    from dmvsnet_b200 import ops
    g = load_golden("warp_edge")
    feats = [cuda(g["feat%d" % i]) for i in range(3)]
    rt = cuda(ops.relative_projections(g["proj"]))
    got = ops.warp_corr(feats, rt, cuda(g["hyp"]))
    assert rel_linf(got, g["cost"]) < 2e-6
    assert float(got[:, :, 0, :4].abs().max()) == 0.0  # behind-camera rows sample outside -> exact zeros


@pytest.mark.parametrize("c,d,n,b,h,w", [(32, 48, 4, 1, 37, 50), (16, 32, 5, 2, 24, 40), (8, 8, 3, 1, 64, 97), (8, 4, 7, 1, 40, 33),
                                          (32, 4, 2, 2, 16, 24), (16, 5, 3, 1, 9, 130)])
def test_warp_corr_vs_oracle(c, d, n, b, h, w, w1_layout):
    from dmvsnet_b200 import ops, synthetic as syn
    g = torch.Generator().manual_seed(c * 1000 + d)
    feats = [torch.randn(b, c, h, w, generator=g) for _ in range(n)]
    proj = syn.make_proj_matrices(max(h, 8) * 4, max(w, 8) * 4, n, b, num_stages=1)["stage1"]
    hyp = 425 + 500 * torch.rand(b, d, h, w, generator=g)
    want = O.warp_corr(feats, proj, hyp)
    got = ops.warp_corr([cuda(f) for f in feats], cuda(ops.relative_projections(proj)), cuda(hyp))
    assert rel_linf(got, want) < 2e-6, rel_linf(got, want)


def test_warp_corr_channel_sliced_views_and_plane_sharding(w1_layout):
    """FeatureNet hands out channel-sliced views ([B,2C,h,w].split); depth shards must tile the full result bit-exactly."""
    from dmvsnet_b200 import ops, synthetic as syn
    g = torch.Generator().manual_seed(5)
    b, c, h, w, d, n = 2, 16, 20, 36, 12, 3
    both = [cuda(torch.randn(b, 2 * c, h, w, generator=g)) for _ in range(n)]
    proj = syn.make_proj_matrices(h * 4, w * 4, n, b, num_stages=1)["stage1"]
    rt = cuda(ops.relative_projections(proj))
    hyp = cuda(425 + 500 * torch.rand(b, d, h, w, generator=g))
    for half in (0, 1):
        views = [t.split([c, c], 1)[half] for t in both]
        full = ops.warp_corr(views, rt, hyp)
        dense = ops.warp_corr([v.contiguous() for v in views], rt, hyp)
        assert torch.equal(full, dense)
        want = O.warp_corr([v.cpu().contiguous() for v in views], proj, hyp.cpu())
        assert rel_linf(full, want) < 2e-6
    sharded = torch.full_like(full, float("nan"))
    for lo, hi in ((0, 5), (5, 6), (6, 12)):
        ops.warp_corr(views, rt, hyp, d_range=(lo, hi), out=sharded)
    assert torch.equal(sharded, full)


def test_warp_corr_cell_output_matches_fp32_output(w1_layout):
    """want_cells: the cost volume in the conv0 cell layout must hold exactly the fp16 hi/lo split of the fp32 output."""
    from dmvsnet_b200 import ops, synthetic as syn
    g = torch.Generator().manual_seed(21)
    b, c, h, w, d, n = 2, 8, 19, 45, 5, 3
    feats = [cuda(torch.randn(b, c, h, w, generator=g)) for _ in range(n)]
    rt = cuda(ops.relative_projections(syn.make_proj_matrices(h * 4, w * 4, n, b, num_stages=1)["stage1"]))
    hyp = cuda(425 + 500 * torch.rand(b, d, h, w, generator=g))
    cost, cells = ops.warp_corr(feats, rt, hyp, want_cells=True)
    assert torch.equal(cells, _cost_cells_from_volume(cost))
    only_cells = ops.warp_corr(feats, rt, hyp, want_f32=False, want_cells=True)
    assert only_cells[0] is None and torch.equal(only_cells[1], cells)


def test_warp_corr_identity_homography_is_autocorrelation(w1_layout):
    """src == ref and P_src == P_ref: the warp is the identity for every depth, cost = mean_j ref[2j+g]^2."""
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(11)
    b, c, h, w, d = 1, 8, 33, 47, 3
    ref = cuda(torch.randn(b, c, h, w, generator=g))
    rt = torch.zeros(b, 1, 12)
    rt[:, :, 0] = rt[:, :, 4] = rt[:, :, 8] = 1.0
    hyp = cuda(400 + 300 * torch.rand(b, d, h, w, generator=g))
    got = ops.warp_corr([ref, ref], cuda(rt), hyp)
    want = (ref.view(b, c // 2, 2, h, w) ** 2).mean(1).unsqueeze(2).expand(-1, -1, d, -1, -1)
    assert rel_linf(got, want) < 1e-5


def test_warp_corr_channels_last_sources_in_place():
    """Sources that already are channel-last in memory (torch channels_last, incl. channel slices of a [B,2C,h,w] map,
    pixel stride 2C) are gathered in place: same bits as after the library's own repack, and as the dense NHWC copy."""
    from dmvsnet_b200 import ops, synthetic as syn
    g = torch.Generator().manual_seed(9)
    b, c, h, w, d, n = 2, 8, 21, 52, 8, 4
    both = [cuda(torch.randn(b, 2 * c, h, w, generator=g)) for _ in range(n)]
    proj = syn.make_proj_matrices(h * 4, w * 4, n, b, num_stages=1)["stage1"]
    rt = cuda(ops.relative_projections(proj))
    hyp = cuda(425 + 500 * torch.rand(b, d, h, w, generator=g))
    both_cl = [t.contiguous(memory_format=torch.channels_last) for t in both]
    for half in (0, 1):
        views = [t.split([c, c], 1)[half] for t in both]
        views_cl = [t.split([c, c], 1)[half] for t in both_cl]
        assert ops._nhwc_strides(views_cl[1]) == (2 * c, h * w * 2 * c)
        want = O.warp_corr([v.cpu().contiguous() for v in views], proj, hyp.cpu())
        launches = _lib_launches()
        got = ops.warp_corr(views_cl, rt, hyp, layout="nhwc")
        assert _lib_launches() - launches == 1  # no repack kernels
        assert rel_linf(got, want) < 2e-6
        assert torch.equal(got, ops.warp_corr(views, rt, hyp, layout="nhwc"))
        repacked = ops.features_nhwc(views[1])
        assert torch.equal(repacked, views[1]) and ops._nhwc_strides(repacked) == (c, h * w * c)


# ------------------------------------------------------------------------------------------ W1 backward (N2)
@pytest.mark.parametrize("c,d,n,b,h,w", [(32, 6, 3, 1, 37, 50), (16, 9, 4, 2, 24, 40), (8, 8, 3, 1, 64, 97), (8, 3, 7, 1, 40, 33),
                                          (16, 1, 2, 1, 9, 130)])
def test_warp_corr_backward_vs_oracle_autograd(c, d, n, b, h, w):
    """dmvs_warp_corr_backward_f32 against autograd through the oracle (F.grid_sample backward on the CPU): gradients
    w.r.t. the reference and every source feature map.  Tolerance 1e-5 of the largest gradient entry: the sums over
    samples are taken in a different (and, for the scatter, unordered) sequence."""
    from dmvsnet_b200 import ops, synthetic as syn
    g = torch.Generator().manual_seed(c * 100 + d)
    feats = [torch.randn(b, c, h, w, generator=g, requires_grad=True) for _ in range(n)]
    proj = syn.make_proj_matrices(max(h, 8) * 4, max(w, 8) * 4, n, b, num_stages=1)["stage1"]
    hyp = 425 + 500 * torch.rand(b, d, h, w, generator=g)
    gout = torch.randn(b, 2, d, h, w, generator=g)
    O.warp_corr(feats, proj, hyp).backward(gout)
    want = [f.grad for f in feats]
    dev = [cuda(f.detach()).requires_grad_(True) for f in feats]
    rt = cuda(ops.relative_projections(proj))
    cost = ops.warp_corr_autograd(dev, rt, cuda(hyp))
    assert torch.equal(cost.detach(), ops.warp_corr([f.detach() for f in dev], rt, cuda(hyp)))
    cost.backward(cuda(gout))
    for i in range(n):
        assert dev[i].grad.shape == want[i].shape
        assert rel_linf(dev[i].grad, want[i]) < 1e-5, (i, rel_linf(dev[i].grad, want[i]))


def test_warp_corr_backward_layouts_padding_and_partial_grads():
    """Channel-last / channel-sliced inputs give the same gradients as dense NCHW ones; samples in the zero padding carry
    no gradient; only the inputs that require a gradient receive one; the adjoint identity <g, W1(f)> = <dW1^T g, f> / 1
    holds for the bilinear form (W1 is linear in the sources for a fixed reference map)."""
    from dmvsnet_b200 import ops, synthetic as syn
    g = torch.Generator().manual_seed(77)
    b, c, h, w, d, n = 2, 8, 21, 52, 5, 3
    both = [cuda(torch.randn(b, 2 * c, h, w, generator=g)) for _ in range(n)]
    proj = syn.make_proj_matrices(h * 4, w * 4, n, b, num_stages=1)["stage1"]
    rt = cuda(ops.relative_projections(proj))
    hyp = cuda(425 + 500 * torch.rand(b, d, h, w, generator=g))
    gout = cuda(torch.randn(b, 2, d, h, w, generator=g))
    both_cl = [t.contiguous(memory_format=torch.channels_last) for t in both]
    views = [t.split([c, c], 1)[1] for t in both]
    views_cl = [t.split([c, c], 1)[1] for t in both_cl]
    dense = ops.warp_corr_backward([v.contiguous() for v in views], rt, hyp, gout)
    for other in (ops.warp_corr_backward(views, rt, hyp, gout), ops.warp_corr_backward(views_cl, rt, hyp, gout)):
        for a, o in zip(dense, other):
            assert rel_linf(o, a) < 1e-5
    # linear in the sources: <gout, cost> == sum_s <grad_src_s, src_s>, and == <grad_ref, ref>
    cost = ops.warp_corr(views, rt, hyp)
    lhs = float((cost.double() * gout.double()).sum())
    via_src = sum(float((gs.double() * v.double()).sum()) for gs, v in zip(dense[1:], views[1:]))
    via_ref = float((dense[0].double() * views[0].double()).sum())
    scale = float((cost.double() * gout.double()).abs().sum())
    assert abs(lhs - via_src) < 1e-5 * scale and abs(lhs - via_ref) < 1e-5 * scale
    # a translation that throws every sample far outside the source images (zero padding) -> exact zero gradients
    far = rt.clone()
    far[:, :, 9] = 1e7
    assert float(ops.warp_corr(views, far, hyp).abs().max()) == 0.0
    zero = ops.warp_corr_backward(views, far, hyp, gout)
    assert all(float(z.abs().max()) == 0.0 for z in zero)
    # partial requires_grad
    leaf = [v.contiguous().requires_grad_(i != 1) for i, v in enumerate(views)]
    ops.warp_corr_autograd(leaf, rt, hyp).backward(gout)
    assert leaf[1].grad is None and leaf[0].grad is not None and leaf[2].grad is not None
    assert rel_linf(leaf[2].grad, dense[2]) < 1e-5


def test_cost_agg_module_is_differentiable_like_the_reference():
    """``CostAgg.forward`` (mvsnet.py:111) with feature maps that require a gradient records the native backward; d loss / d features
    equals autograd through the oracle.  Without requires_grad (or under no_grad) it stays the plain forward."""
    from dmvsnet_b200 import CostAgg, synthetic as syn
    g = torch.Generator().manual_seed(3)
    b, c, h, w, d, n = 1, 16, 18, 30, 6, 3
    feats = [torch.randn(b, c, h, w, generator=g) for _ in range(n)]
    proj = syn.make_proj_matrices(h * 4, w * 4, n, b, num_stages=1)["stage1"]
    hyp = 425 + 500 * torch.rand(b, d, h, w, generator=g)
    leaf = [f.clone().requires_grad_(True) for f in feats]
    (O.warp_corr(leaf, proj, hyp) ** 2).sum().backward()
    dev = [cuda(f).requires_grad_(True) for f in feats]
    agg = CostAgg("variance")
    cost = agg(dev, proj, cuda(hyp), 0)
    assert cost.requires_grad
    (cost ** 2).sum().backward()
    for a, o in zip(dev, leaf):
        assert rel_linf(a.grad, o.grad) < 2e-5
    with torch.no_grad():
        assert not agg(dev, proj, cuda(hyp), 0).requires_grad
    assert not agg([f.detach() for f in dev], proj, cuda(hyp), 0).requires_grad


@pytest.mark.parametrize("c,d,h,w", [(8, 8, 1184, 1600), (32, 4, 296, 400)])
def test_warp_corr_backward_full_size_adjoint_identity(c, d, h, w):
    """DTU full-size W1 passes (N = 5): W1 is bilinear in (reference map, source maps), so <g, W1(ref, src)> = <dW1/dref^T g, ref>
    = sum_s <dW1/dsrc_s^T g, src_s> - a size-independent check of the backward kernel against the forward one."""
    from dmvsnet_b200 import ops, synthetic as syn
    g = torch.Generator().manual_seed(c + d)
    n = 5
    feats = [cuda(torch.randn(1, c, h, w, generator=g)) for _ in range(n)]
    scale = 1600 // w
    proj = syn.make_proj_matrices(1184, 1600, n, 1, num_stages=3)["stage%d" % {4: 1, 2: 2, 1: 3}[scale]]
    rt = cuda(ops.relative_projections(proj))
    hyp = cuda(425 + 500 * torch.rand(1, d, h, w, generator=g))
    gout = cuda(torch.randn(1, 2, d, h, w, generator=g))
    cost = ops.warp_corr(feats, rt, hyp)
    grads = ops.warp_corr_backward(feats, rt, hyp, gout)
    lhs = float((cost.double() * gout.double()).sum())
    norm = float((cost.double() * gout.double()).abs().sum())
    via_ref = float((grads[0].double() * feats[0].double()).sum())
    via_src = sum(float((gs.double() * f.double()).sum()) for gs, f in zip(grads[1:], feats[1:]))
    assert abs(lhs - via_ref) < 1e-6 * norm and abs(lhs - via_src) < 1e-6 * norm, (lhs, via_ref, via_src, norm)


def _lib_launches():
    from dmvsnet_b200 import _native
    return _native.launch_count()


# ------------------------------------------------------------------------------------------ N1 FeatureNet
@pytest.mark.parametrize("b,n,h,w", [(1, 3, 64, 96), (2, 2, 96, 160)])
def test_feature_net_native_vs_oracle(b, n, h, w):
    """dmvs_conv2d_f32 layer chain (fp32 direct convolutions, channel-last outputs) against the oracle's FeatureNet
    (torch CPU conv2d, module.py:274-340): summation order only -> 1e-5 of max|feature|; the cuDNN engine as well."""
    from dmvsnet_b200 import MVSNet, synthetic as syn
    net = MVSNet([8, 8, 8], [4, 2, 1])
    state = syn.randomise_regnet_state(net.state_dict(), seed=3)
    net.load_state_dict(state)
    net = net.to(DEV).eval()
    imgs = syn.make_images(h, w, n, b, seed=4)
    fp = O._sub(state, "feature.")
    want = [O.feature_net(imgs[:, v], fp) for v in range(n)]
    for engine in ("native", "cudnn"):
        net.feature.engine = engine
        launches = _lib_launches()
        with torch.no_grad():
            got = net.extract_features(cuda(imgs))
        # 13 layers + the pixel-unshuffle of c1 for the second 5x5 stride-2 layer; the cuDNN engine launches nothing of ours
        # ... + one fp16 conversion per map for the source views (w1_precision = "fp16", the default); the native tensor heads
        # (out2 / out3) write the fp16 copies of their four maps themselves, so only the two stage-1 maps are converted
        assert net.w1_precision == "fp16"
        assert _lib_launches() - launches == ((14 + 2) if engine == "native" else 6)
        assert torch.equal(got[1]["stage2_h16"].float(), got[1]["stage2"].half().float())
        assert "stage2_h16" not in net.feature(cuda(imgs[:, 0]))  # FeatureNet.forward itself returns the reference's keys only
        assert got[1]["stage2_h16"].shape == got[1]["stage2"].shape and "stage2_h16" not in got[0]
        assert torch.equal(got[1]["stage3_c_h16"].float(), got[1]["stage3_c"].half().float())
        if engine == "native":  # ... and with the 3x3 heads on the fp32 kernels instead of the tensor cores
            net.feature.tensor_heads = False
            with torch.no_grad():
                got32 = net.extract_features(cuda(imgs))
            net.feature.tensor_heads = True
            for v in range(n):
                for key, ref in want[v].items():
                    assert rel_linf(got32[v][key], ref) < 1e-5, ("fp32 heads", v, key)
        for v in range(n):
            for key, ref in want[v].items():
                assert got[v][key].shape == ref.shape
                assert rel_linf(got[v][key], ref) < 1e-5, (engine, v, key, rel_linf(got[v][key], ref))
    net.feature.engine = "native"
    with torch.no_grad():
        got = net.extract_features(cuda(imgs))
    from dmvsnet_b200 import ops
    assert ops._nhwc_strides(got[1]["stage3"]) is not None and ops._nhwc_strides(got[0]["stage2_c"]) is not None
    assert ops._nhwc_strides(got[1]["stage1"]) is not None


def test_conv2d_rejects_foreign_shapes():
    from dmvsnet_b200 import ops
    layer = ops.PackedConv2d(torch.zeros(8, 4, 3, 3, device=DEV))
    with pytest.raises(RuntimeError, match="not a FeatureNet layer shape"):
        ops.conv2d(torch.zeros(1, 4, 8, 8, device=DEV), layer)


# ------------------------------------------------------------------------------------------ R1 single layers
def _torch_block(x, w, bn, stride, transposed, relu, skip):
    import torch.nn.functional as F
    if w.dim() == 4:
        xs = x.squeeze(2)
        y = (F.conv_transpose2d(xs, w, None, stride=2, padding=1, output_padding=1) if transposed
             else F.conv2d(xs, w, None, stride=stride, padding=1)).unsqueeze(2)
    else:
        y = (F.conv_transpose3d(x, w, None, stride=2, padding=1, output_padding=1) if transposed
             else F.conv3d(x, w, None, stride=stride, padding=1))
    if bn is not None:
        y = F.batch_norm(y, bn[2], bn[3], bn[0], bn[1], False, 0.1, 1e-5)
    if relu:
        y = F.relu(y)
    return y if skip is None else y + skip


@pytest.mark.parametrize("cin,cout,dims,stride,transposed,two_d", [
    (2, 8, (8, 16, 24), 1, False, False), (8, 16, (8, 16, 24), 2, False, False), (16, 16, (4, 9, 35), 1, False, False),
    (16, 32, (4, 8, 12), 2, False, False), (32, 32, (2, 7, 33), 1, False, False), (32, 64, (2, 8, 8), 2, False, False),
    (64, 64, (1, 5, 6), 1, False, False), (64, 32, (1, 4, 6), 2, True, False), (32, 16, (2, 8, 12), 2, True, False),
    (16, 8, (4, 6, 37), 2, True, False), (8, 2, (8, 16, 24), 1, False, False), (32, 64, (1, 8, 12), 2, False, True),
    (64, 64, (1, 6, 7), 1, False, True), (64, 32, (1, 4, 6), 2, True, True), (8, 16, (5, 7, 9), 2, False, False),
    (16, 16, (3, 5, 70), 1, False, False), (16, 16, (5, 37, 50), 1, False, False), (32, 32, (3, 19, 27), 1, False, False),
    (64, 64, (2, 17, 9), 1, False, False), (8, 2, (7, 33, 41), 1, False, False), (2, 8, (6, 35, 29), 1, False, False),
    (8, 16, (8, 34, 50), 2, False, False), (16, 32, (4, 37, 21), 2, False, False), (32, 64, (3, 18, 25), 2, False, False),
    (16, 8, (3, 19, 13), 2, True, False), (32, 16, (2, 17, 11), 2, True, False), (64, 32, (1, 9, 10), 2, True, False)])
@pytest.mark.parametrize("engine", ["fp32"])  # single layers on fp32 NCDHW tensors: the exact kernels (tensor engine: the ch16 tests below)
def test_conv_layer_vs_torch(cin, cout, dims, stride, transposed, two_d, engine):
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(cin * 100 + cout)
    b = 2
    x = torch.randn(b, cin, *dims, generator=g)
    ksz = (3, 3) if two_d else (3, 3, 3)
    w = torch.randn(*((cin, cout) if transposed else (cout, cin)), *ksz, generator=g) * (2.0 / (cin * 9 * (1 if two_d else 3))) ** 0.5
    has_bn = cout != 2
    bn = (0.8 + 0.4 * torch.rand(cout, generator=g), 0.1 * torch.randn(cout, generator=g), 0.2 * torch.randn(cout, generator=g),
          0.5 + torch.rand(cout, generator=g)) if has_bn else None
    want = _torch_block(x, w, bn, stride, transposed, has_bn, None)
    skip = torch.randn(want.shape, generator=g) if transposed else None
    if skip is not None:
        want = want + skip
    layer = ops.PackedLayer(cuda(w), transposed, tuple(cuda(t) for t in bn) if bn else None)
    got = ops.conv3d(cuda(x), layer, stride=stride, relu=has_bn, skip=cuda(skip) if skip is not None else None, engine=engine)
    assert tuple(got.shape) == tuple(want.shape)
    assert rel_linf(got, want) < 1e-5, rel_linf(got, want)


def _cost_cells_from_volume(vol):
    """Pack an fp32 [B,2,D,H,W] volume into DMVS_FMT_COST2 cells with torch (the layout dmvs_warp_corr_f32 emits)."""
    b, _, d, h, w = vol.shape
    hi = vol.to(torch.float16)
    lo = (vol - hi.float()).to(torch.float16)
    vox = torch.stack((hi[:, 0], hi[:, 1], lo[:, 0], lo[:, 1]), dim=-1)           # [B,D,H,W,4] halfs of one voxel
    cells = torch.zeros(b, d, h, w + 1, 8, dtype=torch.float16, device=vol.device)
    cells[:, :, :, 1:, 0:4] = vox       # first half of cell x+1 = voxel x ... i.e. cell x = [voxel x-1 | voxel x]
    cells[:, :, :, :w, 4:8] = vox
    return cells.view(torch.int32).reshape(b, d, h, w + 1, 4)


@pytest.mark.parametrize("cfg", [
    (2, 8, (6, 35, 24), 1, False), (8, 16, (8, 34, 48), 2, False), (16, 16, (5, 37, 50), 1, False), (16, 32, (4, 37, 22), 2, False),
    (32, 32, (3, 19, 27), 1, False), (32, 64, (3, 18, 26), 2, False), (64, 64, (2, 17, 9), 1, False), (64, 32, (1, 9, 10), 2, True),
    (32, 16, (2, 17, 11), 2, True), (16, 8, (3, 19, 13), 2, True), (8, 2, (7, 33, 41), 1, False), (8, 16, (1, 8, 8), 2, False),
    (16, 8, (1, 40, 100), 2, True),
    # depth-tap-folded kernels (csrc/conv_kf.cu): one and two planes, ranges that cut columns, odd plane counts
    (16, 16, (1, 18, 26), 1, False), (16, 16, (2, 37, 50), 1, False), (16, 16, (3, 148, 200), 1, False), (8, 2, (4, 50, 60), 1, False),
    (32, 32, (1, 20, 30), 1, False), (32, 32, (2, 37, 50), 1, False), (32, 32, (5, 74, 100), 1, False),
    (8, 2, (1, 20, 30), 1, False), (2, 8, (4, 50, 61), 1, False), (2, 8, (1, 20, 31), 1, False), (2, 8, (9, 150, 200), 1, False),
    # many tiles per persistent CTA: the stage ring and both accumulator sets wrap several times
    (8, 2, (24, 160, 200), 1, False), (2, 8, (16, 160, 200), 1, False), (16, 16, (12, 160, 200), 1, False),
    (8, 16, (12, 160, 200), 2, False), (16, 8, (6, 80, 100), 2, True), (32, 32, (8, 80, 104), 1, False),
    (32, 64, (4, 80, 104), 2, False), (64, 64, (3, 40, 56), 1, False), (64, 32, (2, 40, 56), 2, True), (32, 16, (4, 40, 56), 2, True),
    # the refine net's 2-D bottleneck (kd = 1)
    (32, 64, (1, 18, 26), 2, False, True), (64, 64, (1, 17, 9), 1, False, True), (64, 32, (1, 9, 10), 2, True, True),
    (32, 64, (1, 148, 200), 2, False, True), (64, 64, (1, 74, 100), 1, False, True), (64, 32, (1, 74, 100), 2, True, True)])
def test_conv_layer_ch16_vs_torch(cfg):
    """The TMA-fed persistent tcgen05 kernels on the CH16 / CH16P cell layouts, one layer at a time, vs torch fp32."""
    from dmvsnet_b200 import ops
    cin, cout, dims, stride, transposed = cfg[:5]
    two_d = len(cfg) > 5 and cfg[5]
    g = torch.Generator().manual_seed(cin * 100 + cout + 7)
    b = 2 if dims[1] < 80 else 1
    x = torch.randn(b, cin, *dims, generator=g)
    ksz = (3, 3) if two_d else (3, 3, 3)
    w = torch.randn(*((cin, cout) if transposed else (cout, cin)), *ksz, generator=g) * (2.0 / (cin * 9 * (1 if two_d else 3))) ** 0.5
    has_bn = cout != 2
    bn = (0.8 + 0.4 * torch.rand(cout, generator=g), 0.1 * torch.randn(cout, generator=g), 0.2 * torch.randn(cout, generator=g),
          0.5 + torch.rand(cout, generator=g)) if has_bn else None
    want = _torch_block(x, w, bn, stride, transposed, has_bn, None)
    skip = torch.randn(want.shape, generator=g) if transposed else None
    if skip is not None:
        want = want + skip
    layer = ops.PackedLayer(cuda(w), transposed, tuple(cuda(t) for t in bn) if bn else None)
    xin = cuda(x) if cin == 2 else ops.to_ch16(cuda(x), parity_split=(stride == 2 and not transposed))
    if cin == 2:  # the same layer fed by W1's cell-format cost volume (identity warp of a 2-channel "feature" = the volume itself)
        cells = _cost_cells_from_volume(cuda(x))
        via_cells = ops.from_ch16(ops.conv3d_ch16(cells, layer, relu=True, out_fmt="ch16", in_cells=True), cout)
        assert rel_linf(via_cells, want) < 1e-5, rel_linf(via_cells, want)
    if cin != 2:  # the converters round-trip to 2^-22
        back = ops.from_ch16(xin, cin, parity_split=(stride == 2 and not transposed))
        assert rel_linf(back, x) < 1e-6
    sk = ops.to_ch16(cuda(skip), parity_split=True) if skip is not None else None
    if cout == 2:
        got = ops.conv3d_ch16(xin, layer, stride=stride, relu=False, out_fmt="f32")
    else:
        for fmt in (["ch16"] if (transposed or two_d) else ["ch16", "ch16p"]):
            if fmt == "ch16p" and want.shape[-1] % 2:
                continue
            cells = ops.conv3d_ch16(xin, layer, stride=stride, relu=True, skip=sk, out_fmt=fmt)
            got = ops.from_ch16(cells, cout, parity_split=(fmt == "ch16p"))
            assert rel_linf(got, want) < 1e-5, (fmt, rel_linf(got, want))
    assert tuple(got.shape) == tuple(want.shape)
    assert rel_linf(got, want) < 1e-5, rel_linf(got, want)


@pytest.mark.parametrize("cfg", [(8, 2, (4, 50, 60), 1, False), (8, 2, (24, 160, 200), 1, False), (2, 8, (4, 50, 61), 1, False),
                                 (2, 8, (9, 150, 200), 1, False), (16, 16, (3, 148, 200), 1, False)])
def test_conv_depth_tap_folded_kernels_all_kinds(cfg, native_lib):
    """csrc/conv_kf.cu serves conv2 by default; its conv0 / prob kinds (knob kf = 2) and the wide-tile prob kernel (knob kf_wide;
    measured no faster than the per-tap kernels) stay correct."""
    try:
        assert native_lib.dmvs_debug_set(b"kf", 2) == 0
        for mw in (0, 1):  # issuing-thread variants of every kind
            assert native_lib.dmvs_debug_set(b"kf_mw", mw) == 0
            test_conv_layer_ch16_vs_torch(cfg)
            if cfg[:2] == (8, 2):
                assert native_lib.dmvs_debug_set(b"kf_wide", 1) == 0
                test_conv_layer_ch16_vs_torch(cfg)
                assert native_lib.dmvs_debug_set(b"kf_wide", 0) == 0
    finally:
        native_lib.dmvs_debug_set(b"kf", 1)
        native_lib.dmvs_debug_set(b"kf_wide", 0)
        native_lib.dmvs_debug_set(b"kf_mw", 0)


@pytest.mark.parametrize("engine", ["fp32", "tensor"])
@pytest.mark.parametrize("refine,d,h,w,b", [(False, 8, 16, 24, 1), (False, 16, 8, 40, 2), (True, 4, 16, 24, 2), (True, 4, 40, 8, 1)])
def test_regnet_vs_oracle(refine, d, h, w, b, engine):
    from dmvsnet_b200 import MVSNet, ops, synthetic as syn
    net = MVSNet([8, 8, 8], [4, 2, 1])
    state = syn.randomise_regnet_state(net.state_dict(), seed=2)
    net.load_state_dict(state)
    net = net.to(DEV).eval()
    mod = (net.cost_regularization_refine if refine else net.cost_regularization)[1]
    prefix = "cost_regularization%s.1." % ("_refine" if refine else "")
    g = torch.Generator().manual_seed(d + h)
    x = torch.randn(b, 2, d, h, w, generator=g)
    with torch.no_grad():
        want = O.regnet(x, O._sub(state, prefix), refine=refine)
        ops.DEFAULT_ENGINE, saved = engine, ops.DEFAULT_ENGINE
        try:
            got = mod(cuda(x))
        finally:
            ops.DEFAULT_ENGINE = saved
        # the single-branch layer-by-layer route must agree with the fused driver
        small = mod.cosR_small(cuda(x))
    assert rel_linf(got, want) < 1e-4, rel_linf(got, want)
    assert rel_linf(small, want[:, :2]) < 1e-4


@pytest.mark.parametrize("refine,d,h,w", [(False, 16, 40, 56), (True, 4, 48, 40)])
def test_regnet_programmatic_dependent_launch_is_bit_identical(refine, d, h, w, native_lib):
    """The tensor convs are launched as programmatic dependents (their prologue overlaps the predecessor's tail): the logits
    must carry the same bits as with plain stream-ordered launches, run after run."""
    from dmvsnet_b200 import MVSNet, synthetic as syn
    net = MVSNet([8, 8, 8], [4, 2, 1])
    net.load_state_dict(syn.randomise_regnet_state(net.state_dict(), seed=4))
    net = net.to(DEV).eval()
    mod = (net.cost_regularization_refine if refine else net.cost_regularization)[2]
    x = cuda(torch.randn(1, 2, d, h, w, generator=torch.Generator().manual_seed(h)))
    try:
        with torch.no_grad():
            assert native_lib.dmvs_debug_set(b"tc2_pdl", 0) == 0
            plain = mod(x).clone()
            assert native_lib.dmvs_debug_set(b"tc2_pdl", 1) == 0
            for _ in range(5):
                assert torch.equal(mod(x), plain)
    finally:
        native_lib.dmvs_debug_set(b"tc2_pdl", 1)


# ------------------------------------------------------------------------------------------ E1 / E2 / S1
@pytest.mark.parametrize("d,h,w,b", [(48, 9, 37, 1), (8, 16, 33, 2), (5, 7, 5, 1)])
def test_depth_head_vs_oracle(d, h, w, b):
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(d)
    logits = 6 * torch.randn(b, 4, d, h, w, generator=g)
    hyp = 425 + 500 * torch.rand(b, d, h, w, generator=g).sort(1)[0]
    interval = torch.tensor(10.77)
    want = O.depth_head(logits, hyp, interval)
    prob, d4, hyp_c, conf = ops.depth_head(cuda(logits), cuda(hyp), cuda(interval))
    assert rel_linf(prob, want["prob_volume"]) < 1e-5
    assert rel_linf(d4, want["depth_sub_plus"]) < 1e-6
    assert rel_linf(hyp_c, want["depth_values_c"]) < 1e-5
    assert float((conf.cpu() - want["photometric_confidence"]).abs().max()) < 1e-4
    assert ops.depth_head(cuda(logits), cuda(hyp), 10.77, want_prob=False)[0] is None


@pytest.mark.parametrize("h,w,b", [(9, 37, 1), (16, 33, 2)])
def test_refine_head_vs_oracle(h, w, b):
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(h)
    logits = 3 * torch.randn(b, 4, 4, h, w, generator=g)
    hyp_c = 425 + 500 * torch.rand(b, 4, h, w, generator=g)
    want = O.refine_head(logits, hyp_c, torch.tensor(5.0))
    depth, conf, d4 = ops.refine_head(cuda(logits), cuda(hyp_c), 5.0)
    assert rel_linf(d4, want["depth_sub_plus_refine"]) < 1e-6
    assert rel_linf(depth, want["depth"]) < 1e-6
    assert float((conf.cpu() - want["photometric_confidence_refine"]).abs().max()) < 1e-4


@pytest.mark.parametrize("inverse", [False, True])
def test_hypotheses_first_vs_oracle(inverse):
    from dmvsnet_b200 import ops, synthetic as syn
    dv = syn.make_depth_values(2, 192, inverse=inverse)
    dv[1] = dv[1] * 1.1
    want, wi = O.depth_hypotheses(dv, 48, None, (9, 21), inverse)
    got, gi = ops.hypotheses_first(cuda(dv), 48, (9, 21), inverse)
    assert rel_linf(got, want) < 1e-6, rel_linf(got, want)
    assert abs(float(gi) - float(wi)) < 1e-5 * float(wi)


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("d", [32, 8])
def test_hypotheses_next_vs_oracle(inverse, d):
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(d)
    last = 450 + 400 * torch.rand(2, 11, 14, generator=g)
    ip = torch.tensor(2.65 * 2)
    lo_res, wi = O.depth_hypotheses(last, d, ip, None, inverse)
    want = O.upsample_hypotheses(lo_res, (22, 28))
    got, gi = ops.hypotheses_next(cuda(last), d, cuda(ip), (22, 28), inverse)
    assert rel_linf(got, want) < 1e-6, rel_linf(got, want)
    assert abs(float(gi) - float(wi)) < 1e-6 * float(wi)
    same, _ = ops.hypotheses_next(cuda(last), d, cuda(ip), None, inverse)
    assert rel_linf(same, lo_res) < 1e-6


# ------------------------------------------------------------------------------------------ whole cascade vs reference fixtures
# seam -> (tolerance, how): "rel" = per-element relative (depths), "lin" = relative to max|want| (volumes), "abs"
CASCADE_SEAMS = [("depth_values", 1e-5, "rel"), ("cost", 2e-5, "lin"), ("logits", 2e-4, "lin"), ("depth_sub_plus", 1e-3, "rel"),
                 ("depth_values_c", 1e-3, "rel"), ("photometric_confidence", 1e-2, "abs"), ("cost_c", 1e-3, "lin"),
                 ("logits_c", 2e-3, "lin"), ("depth_sub_plus_refine", 1e-3, "rel"), ("depth", 1e-3, "rel"),
                 ("photometric_confidence_refine", 1e-2, "abs")]


@pytest.mark.parametrize("name", sorted(CASES))
def test_cascade_against_reference_fixture(name):
    """All three stages through MVSNet.cascade / MVSNet.forward against the reference's own outputs, seam by seam.
    The contract (BASELINE.json north_star) is 1e-3 relative on the regressed depth; the table shows what is achieved."""
    from dmvsnet_b200 import MVSNet
    case = CASES[name]
    gold = load_golden(name)
    inp = case_inputs(case)
    net = MVSNet(case["ndepths"], case["ratios"], inverse_depth=case["inverse"])
    net.load_state_dict(case_state(case))
    net = net.to(DEV).eval()
    net.w1_precision = "fp32"  # the exact W1 kernels: these tests assert fp32-level seams (fp16 mode: test_gpu_h16.py)
    with torch.no_grad():
        if "features" in inp:
            feats = [{k: cuda(v) for k, v in f.items()} for f in inp["features"]]
        else:
            feats = net.extract_features(cuda(inp["imgs"]))
        rts = [gold["s%d_rt" % (s + 1)] for s in range(len(case["ndepths"]))]
        out = net.cascade(feats, inp["proj"], cuda(inp["depth_values"]), (case["H"], case["W"]), keep_seams=True, rts=rts)
        # second pass with the homographies recomputed on this host: only the 1e-3 depth contract is asserted
        own = net.cascade(feats, inp["proj"], cuda(inp["depth_values"]), (case["H"], case["W"]))
    last = "s%d_depth" % len(case["ndepths"])
    own_err = float(((own["depth"].cpu() - gold[last]).abs() / gold[last].abs().clamp_min(1.0)).max())
    assert own_err < 1e-3, "host-computed homographies: final depth rel err %.2e" % own_err
    report, bad = ["final depth with host-computed H: rel err %.2e" % own_err], []
    for s in range(len(case["ndepths"])):
        st = out["stage%d" % (s + 1)]
        for seam, tol, how in CASCADE_SEAMS:
            if s > 0 and seam in ("cost", "depth_values"):
                tol = 2e-4  # later stages inherit ~1e-5 relative differences in the regressed depth through the hypotheses
            got = (st["_" + seam] if "_" + seam in st else st[seam]).cpu()
            want = gold["s%d_%s" % (s + 1, seam)]
            diff = (got - want).abs()
            err = float({"rel": (diff / want.abs().clamp_min(1.0)).max(), "lin": diff.max() / want.abs().max(), "abs": diff.max()}[how])
            report.append("stage%d %-30s %s err %.2e (tol %.0e)" % (s + 1, seam, how, err, tol))
            if not err < tol:
                bad.append(report[-1])
    print("\n".join(report))
    assert not bad, "\n".join(["seams out of tolerance:"] + bad + ["all seams:"] + report)
    for k in ("depth", "photometric_confidence", "prob_volume", "depth_values_c", "stage1"):
        assert k in out


def test_forward_from_images_matches_fixture():
    from dmvsnet_b200 import MVSNet
    case = CASES["cfg1_full"]
    gold = load_golden("cfg1_full")
    inp = case_inputs(case)
    net = MVSNet(case["ndepths"], case["ratios"], inverse_depth=case["inverse"])
    net.load_state_dict(case_state(case))
    net = net.to(DEV).eval()
    net.w1_precision = "fp32"  # the exact W1 kernels: these tests assert fp32-level seams (fp16 mode: test_gpu_h16.py)
    with torch.no_grad():
        out = net(cuda(inp["imgs"]), {k: cuda(v) for k, v in inp["proj"].items()}, cuda(inp["depth_values"]))
    err = float(((out["depth"].cpu() - gold["s1_depth"]).abs() / gold["s1_depth"].abs().clamp_min(1.0)).max())
    assert err < 1e-3, err
    assert out["prob_volume"].shape == (1, 4, 48, 32, 40) and out["interval"].dim() == 0


def test_full_three_stage_forward_from_images_vs_oracle():
    """imgs -> native FeatureNet (channel-last features, tensor-core 3x3 layers) -> 3-stage cascade, B = 2, against the oracle's
    mvsnet_forward on the same host (same homographies).  North-star tolerance: 1e-3 relative on the depth map of every stage
    (measured 5e-5)."""
    from dmvsnet_b200 import MVSNet, synthetic as syn
    b, n, h, w, nd, ratios = 2, 3, 64, 96, [16, 8, 8], [4, 2, 1]
    net = MVSNet(nd, ratios, inverse_depth=True)
    state = syn.randomise_regnet_state(net.state_dict(), seed=5)
    net.load_state_dict(state)
    net = net.to(DEV).eval()
    net.w1_precision = "fp32"
    imgs = syn.make_images(h, w, n, b, seed=6)
    proj = syn.make_proj_matrices(h, w, n, b, num_stages=3)
    dv = syn.make_depth_values(b, 192, inverse=True)
    with torch.no_grad():
        got = net(cuda(imgs), proj, cuda(dv))
        want = O.mvsnet_forward(imgs, proj, dv, state, nd, ratios, True)
    for stage in ("stage1", "stage2", "stage3"):
        err = float(((got[stage]["depth"].cpu() - want[stage]["depth"]).abs() / want[stage]["depth"].abs()).max())
        assert err < 1e-3, (stage, err)
    assert torch.equal(got["depth"], got["stage3"]["depth"])
    assert float((got["photometric_confidence"].cpu() - want["photometric_confidence"]).abs().max()) < 1e-2
    assert set(want.keys()) <= set(got.keys())


def test_infer_from_host_buffers():
    from dmvsnet_b200 import MVSNet
    case = CASES["cfg1_full"]
    gold = load_golden("cfg1_full")
    inp = case_inputs(case)
    net = MVSNet(case["ndepths"], case["ratios"], inverse_depth=case["inverse"])
    net.load_state_dict(case_state(case))
    net = net.to(DEV).eval()
    net.w1_precision = "fp32"  # the exact W1 kernels: these tests assert fp32-level seams (fp16 mode: test_gpu_h16.py)
    host = net.infer(inp["imgs"], inp["proj"], inp["depth_values"])
    assert not host["depth"].is_cuda
    err = float(((host["depth"] - gold["s1_depth"]).abs() / gold["s1_depth"].abs().clamp_min(1.0)).max())
    assert err < 1e-3, err


def test_infer_many_pipeline_matches_infer():
    """The three-stage pipeline (upload k+1 | compute k | download k-1) returns, in order, exactly what infer() returns."""
    from dmvsnet_b200 import MVSNet
    case = CASES["cfg1_full"]
    inp = case_inputs(case)
    net = MVSNet(case["ndepths"], case["ratios"], inverse_depth=case["inverse"])
    net.load_state_dict(case_state(case))
    net = net.to(DEV).eval()
    net.w1_precision = "fp32"  # the exact W1 kernels: these tests assert fp32-level seams (fp16 mode: test_gpu_h16.py)
    items = [((inp["imgs"] * s).clamp(0, 1), inp["proj"], inp["depth_values"]) for s in (1.0, 0.7, 0.4, 0.9)]
    want = [net.infer(*it) for it in items]
    got = list(net.infer_many(items))
    assert len(got) == len(want)
    for g, w_ in zip(got, want):
        assert torch.equal(g["depth"], w_["depth"]) and torch.equal(g["photometric_confidence"], w_["photometric_confidence"])
    assert not torch.equal(want[0]["depth"], want[2]["depth"])
    assert list(net.infer_many([])) == []


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_size_dtu_stage3_properties(w1_layout):
    """BASELINE config 2 at full stage-3 size (1184x1600, C=8, N=5): size-independent properties instead of the oracle."""
    from dmvsnet_b200 import ops, synthetic as syn
    h, w, c, n, d = 1184, 1600, 8, 5, 8
    g = torch.Generator().manual_seed(0)
    feats = [cuda(torch.randn(1, c, h, w, generator=g)) for _ in range(n)]
    proj = syn.make_proj_matrices(h, w, n, 1)["stage3"]
    rt = cuda(ops.relative_projections(proj))
    hyp = cuda(425 + 500 * torch.rand(1, d, h, w, generator=g))
    full = ops.warp_corr(feats, rt, hyp)
    assert torch.isfinite(full).all()
    # linear in the reference features, additive over source views, plane shards tile exactly
    scaled = ops.warp_corr([2.0 * feats[0]] + feats[1:], rt, hyp)
    assert torch.equal(scaled, 2.0 * full)
    parts = sum(ops.warp_corr([feats[0], feats[i + 1]], rt[:, i:i + 1].contiguous(), hyp) for i in range(n - 1))
    assert rel_linf(parts, full) < 1e-6
    shard = torch.empty_like(full)
    for r in range(8):
        ops.warp_corr(feats, rt, hyp, d_range=(r, r + 1), out=shard)
    assert torch.equal(shard, full)


def test_ops_follow_the_tensors_device_not_the_current_one():
    """with the model on a non-current device every native launch (and its TMA descriptors, its per-device shared-
    memory opt-in) must go to the tensors' device.  Needs 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from dmvsnet_b200 import MVSNet, synthetic as syn
    nd, ratios, H, W, n = [16, 8, 8], [4, 2, 1], 128, 160, 3
    net = MVSNet(nd, ratios, inverse_depth=True)
    net.load_state_dict(syn.ridge_regnet_state(net.state_dict(), seed=1))
    proj = syn.make_proj_matrices(H, W, n, 1)
    imgs = syn.make_scene_images(H, W, n, proj["stage3"], seed=1)
    dv = syn.make_depth_values(1, 192, inverse=True)
    with torch.no_grad():
        want = net.to("cuda:0").eval()(imgs.to("cuda:0"), proj, dv.to("cuda:0"))["depth"].cpu()
        torch.cuda.set_device(0)
        got = net.to("cuda:1")(imgs.to("cuda:1"), proj, dv.to("cuda:1"))["depth"]
    assert got.device.index == 1 and torch.equal(got.cpu(), want)


def test_infer_uint8_images_equal_float_images():
    """MVSNet.infer / infer_many take the 8-bit photographs as they are (uint8 host tensor, a quarter of the H2D bytes) and
    scale them on the device: same pixel values as the reference's host-side ``np.array(img, float32) / 255.``
    (datasets/general_eval.py:161), hence the same depth map bit for bit."""
    from dmvsnet_b200 import MVSNet, synthetic as syn
    nd, ratios, H, W, n = [16, 8, 8], [4, 2, 1], 128, 160, 3
    net = MVSNet(nd, ratios, inverse_depth=True)
    net.load_state_dict(syn.ridge_regnet_state(net.state_dict(), seed=1))
    net = net.to(DEV).eval()
    proj = syn.make_proj_matrices(H, W, n, 1)
    u8 = (syn.make_scene_images(H, W, n, proj["stage3"], seed=1) * 255).round().to(torch.uint8)
    f32 = torch.from_numpy(u8.numpy().astype("float32") / 255.0)
    dv = syn.make_depth_values(1, 192, inverse=True)
    want = net.infer(f32, proj, dv)
    got = net.infer(u8, proj, dv)
    assert torch.equal(got["depth"], want["depth"]) and torch.equal(got["photometric_confidence"], want["photometric_confidence"])
    many = list(net.infer_many([(u8, proj, dv), (f32, proj, dv)]))
    assert torch.equal(many[0]["depth"], want["depth"]) and torch.equal(many[1]["depth"], want["depth"])
