"""W1, fp16-staged kernel (dmvs_warp_corr_h16_f32) against the oracle.

Two comparisons per case: (i) against the oracle fed with the SAME fp16-rounded source maps - what is left is the sample
position arithmetic (refined reciprocal instead of IEEE divisions, <= 2.5e-4 px at w = 1600; the test features are white noise, so the cost moves by about as much) and summation order: 5e-4 of max|cost|;
(ii) against the oracle on the unrounded maps - the price of the format: 1e-3 of max|cost|.  The contract (BASELINE.json
north_star) is 1e-3 relative on the regressed depth: test_cascade_h16_* check that through the whole cascade.
"""
import pytest
import torch

from conftest import load_golden, rel_linf
from oracle import dmvs_oracle as O
from test_oracle_golden import _c_warp_corr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _lib(native_lib):
    return native_lib


def cuda(t):
    return t.to(DEV)


def _rounded(feats):
    return [feats[0]] + [f.half().float() for f in feats[1:]]


def test_h16_edge_fixture():
    """Rotated rig, out-of-frustum (zero padding), behind-camera and exact Z == 0 samples (the live reference's outputs)."""
    from dmvsnet_b200 import ops
    g = load_golden("warp_edge")
    feats = [g["feat%d" % i] for i in range(3)]
    rt = ops.relative_projections(g["proj"])
    got = ops.warp_corr([cuda(f) for f in feats], cuda(rt), cuda(g["hyp"]), layout="h16")
    assert rel_linf(got, g["cost"]) < 1e-3
    assert float(got[:, :, 0, :4].abs().max()) == 0.0  # behind-camera rows sample outside -> exact zeros
    want16 = O.warp_corr(_rounded(feats), g["proj"], g["hyp"])
    assert rel_linf(got, want16) < 5e-4, rel_linf(got, want16)


@pytest.mark.parametrize("c,d,n,b,h,w", [(32, 48, 4, 1, 37, 50), (16, 32, 5, 2, 24, 40), (8, 8, 3, 1, 64, 97), (8, 4, 7, 1, 40, 33),
                                          (32, 4, 2, 2, 16, 24), (16, 5, 4, 1, 9, 130), (8, 19, 8, 1, 33, 47), (16, 9, 6, 2, 21, 35)])
@pytest.mark.parametrize("kind", ["rough", "planes"])
def test_h16_vs_oracle(c, d, n, b, h, w, kind):
    """rough: white-noise per-pixel hypotheses (every source takes the direct path); planes: the stage-1 sampler's planes
    (boxes fit, staged path; more sources than ring slots for n > 3 / 5)."""
    from dmvsnet_b200 import ops, synthetic as syn
    g = torch.Generator().manual_seed(c * 1000 + d)
    feats = [torch.randn(b, c, h, w, generator=g) for _ in range(n)]
    proj = syn.make_proj_matrices(max(h, 8) * 4, max(w, 8) * 4, n, b, num_stages=1)["stage1"]
    if kind == "rough":
        hyp = 425 + 500 * torch.rand(b, d, h, w, generator=g)
    else:
        hyp = ops.hypotheses_first(cuda(syn.make_depth_values(b, 192, inverse=True)), d, [h, w], True)[0].cpu()
    want = O.warp_corr(feats, proj, hyp)
    want16 = O.warp_corr(_rounded(feats), proj, hyp)
    rt = cuda(ops.relative_projections(proj))
    got = ops.warp_corr([cuda(f) for f in feats], rt, cuda(hyp), layout="h16")
    assert rel_linf(got, want16) < 5e-4, rel_linf(got, want16)
    assert rel_linf(got, want) < 1e-3, rel_linf(got, want)
    # pre-converted sources, cell output, plane shards
    half = [cuda(feats[0])] + [ops.features_nhwc_f16(cuda(f)) for f in feats[1:]]
    cost, cells = ops.warp_corr(half, rt, cuda(hyp), layout="h16", want_cells=True)
    assert torch.equal(cost, got)
    from test_gpu_parity import _cost_cells_from_volume
    assert torch.equal(cells, _cost_cells_from_volume(cost))
    shard = torch.full_like(got, float("nan"))
    cuts = sorted({0, d // 3, (2 * d) // 3 + 1 if d > 2 else d, d})
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ops.warp_corr(half, rt, cuda(hyp), layout="h16", d_range=(lo, hi), out=shard)
    assert torch.equal(shard, got)


def test_h16_channel_last_inputs_and_converter():
    """The converter accepts NCHW, channel-last and channel-sliced channel-last maps and rounds to nearest even."""
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(4)
    b, c, h, w = 2, 16, 13, 27
    both = cuda(torch.randn(b, 2 * c, h, w, generator=g))
    want = both[:, c:].permute(0, 2, 3, 1).half()
    for t in (both[:, c:], both.contiguous(memory_format=torch.channels_last)[:, c:], both[:, c:].contiguous()):
        got = ops.features_nhwc_f16(t)
        assert got.shape == (b, c, h, w) and torch.equal(got.data, want)


W1_FULL = [("dtu", 1, 1184, 1600, 5, 32, 48, False), ("dtu", 2, 1184, 1600, 5, 16, 32, False), ("dtu", 3, 1184, 1600, 5, 8, 8, False),
           ("dtu", 3, 1184, 1600, 5, 8, 4, True), ("tnt", 3, 1056, 1920, 11, 8, 4, True), ("bmvs", 2, 576, 768, 7, 16, 32, False)]


@pytest.mark.parametrize("cfg,stage,H,W,views,c,d,refine", W1_FULL)
def test_h16_full_size_vs_c_oracle(cfg, stage, H, W, views, c, d, refine, c_oracle):
    from dmvsnet_b200 import ops, synthetic as syn
    from test_gpu_fullsize import _mixed_depth
    scale = 2 ** (3 - stage)
    h, w = H // scale, W // scale
    g = torch.Generator().manual_seed(stage * 100 + c + d + views)
    proj = syn.make_proj_matrices(H, W, views, 1, num_stages=3)["stage%d" % stage]
    rt = ops.relative_projections(proj)
    feats = [torch.randn(1, c, h, w, generator=g) for _ in range(views)]
    dv = syn.make_depth_values(1, 192, inverse=True)
    interval = (dv[0, -1] - dv[0, 0]) / dv.size(1)
    if refine:
        last = _mixed_depth(h, w, g)
        hyp = torch.stack([last + 5.0 * (k - 1.5) for k in range(4)], 1)[:, :d].contiguous()
    elif stage == 1:
        hyp = ops.hypotheses_first(cuda(dv), d, [h, w], True)[0].cpu()
    else:
        ratio = {2: 2.0, 3: 1.0}[stage]
        hyp = ops.hypotheses_next(cuda(_mixed_depth(h // 2, w // 2, g)), d, cuda(ratio * interval), [h, w], True)[0].cpu()
    want16 = _c_warp_corr(c_oracle, _rounded(feats), rt, hyp)
    got = ops.warp_corr([cuda(f) for f in feats], cuda(rt), cuda(hyp), layout="h16")
    assert rel_linf(got, want16) < 5e-4, rel_linf(got, want16)


# ------------------------------------------------------------------------------------------ through the cascade
def test_cascade_h16_reference_fixture():
    """The cascade with fp16-staged W1 against the live reference's outputs (cascade_lin fixture): the cost volumes within
    1e-3 of max|cost|; the App. D random weights regress a white-noise depth map (a peaked softmax at a random plane per
    pixel), so the depth is compared where the conditioning allows it: 99 % of the pixels within the 1e-3 contract."""
    from cases import CASES, case_inputs, case_state
    from dmvsnet_b200 import MVSNet
    case, gold = CASES["cascade_lin"], load_golden("cascade_lin")
    inp = case_inputs(case)
    net = MVSNet(case["ndepths"], case["ratios"], inverse_depth=case["inverse"])
    net.load_state_dict(case_state(case))
    net = net.to(DEV).eval()
    assert net.w1_precision == "fp16"
    feats = [{k: cuda(v) for k, v in f.items()} for f in inp["features"]]
    rts = [gold["s%d_rt" % (s + 1)] for s in range(3)]
    with torch.no_grad():
        out = net.cascade(feats, inp["proj"], cuda(inp["depth_values"]), (case["H"], case["W"]), keep_seams=True, rts=rts)
    assert rel_linf(out["stage1"]["_cost"], gold["s1_cost"]) < 1e-3
    for s in range(3):
        want = gold["s%d_depth" % (s + 1)]
        e = ((out["stage%d" % (s + 1)]["depth"].cpu() - want).abs() / want.abs()).flatten()
        p99 = float(e.kthvalue(int(0.99 * e.numel()))[0])
        print("stage%d depth rel err: max %.2e p99 %.2e mean %.2e" % (s + 1, float(e.max()), p99, float(e.mean())))
        assert p99 < 1e-3


@pytest.mark.parametrize("H,W,views", [(256, 320, 5), (192, 256, 8)])
def test_cascade_h16_conditioned_vs_oracle(H, W, views):
    """Free-running 3-stage cascade, fp16-staged W1, on the conditioned workload (photo-consistent features + ridge-following
    regularisation nets): every stage's depth within the 1e-3 contract of the oracle's fp32 result (max over all pixels);
    the fp16 feature copies attached by add_half_features are used in place."""
    from dmvsnet_b200 import MVSNet, ops, synthetic as syn
    nd, ratios = [48, 32, 8], [4, 2, 1]
    net = MVSNet(nd, ratios, inverse_depth=True)
    state = syn.ridge_regnet_state(net.state_dict(), seed=0)
    net.load_state_dict(state)
    net = net.to(DEV).eval()
    proj = syn.make_proj_matrices(H, W, views, 1, num_stages=3)
    feats = syn.make_scene_features(H, W, views, proj, seed=0)
    dv = syn.make_depth_values(1, 192, inverse=True)
    dfeats = net.add_half_features([{k: cuda(v) for k, v in f.items()} for f in feats])
    assert isinstance(dfeats[1]["stage2_c_h16"], ops.HalfFeatures) and "stage1_h16" not in dfeats[0]
    with torch.no_grad():
        want = O.cascade_forward(feats, proj, dv, state, nd, ratios, True, (H, W))
        out = net.cascade(dfeats, proj, cuda(dv), (H, W))
        net.w1_precision = "fp32"
        exact = net.cascade(dfeats, proj, cuda(dv), (H, W))
    for s in range(3):
        name = "stage%d" % (s + 1)
        e16 = float(((out[name]["depth"].cpu() - want[name]["depth"]).abs() / want[name]["depth"].abs()).max())
        e32 = float(((exact[name]["depth"].cpu() - want[name]["depth"]).abs() / want[name]["depth"].abs()).max())
        print("%s depth rel err vs oracle: fp16-staged W1 %.2e, exact W1 %.2e" % (name, e16, e32))
        assert e16 < 1e-3 and e32 < 1e-3


# ------------------------------------------------------------------------------------------ row bands / single-view sharding
@pytest.mark.parametrize("c,d,h,w,n", [(8, 8, 96, 80, 4), (32, 5, 64, 48, 3)])
def test_h16_row_band_equals_rows_of_the_full_call(c, d, h, w, n):
    """dmvs_warp_corr_h16_f32 on rows [r0, r1) of the reference view (whole source maps, ref_row0 = r0) == the same rows of the
    unsharded call, bit for bit: what the row-band cascade of the single-view sharded mode relies on."""
    from dmvsnet_b200 import ops, synthetic as syn
    g = torch.Generator().manual_seed(c + h)
    feats = [cuda(torch.randn(1, c, h, w, generator=g)) for _ in range(n)]
    proj = syn.make_proj_matrices(h * 4, w * 4, n, 1, num_stages=1)["stage1"]
    rt = cuda(ops.relative_projections(proj))
    hyp = ops.hypotheses_first(cuda(syn.make_depth_values(1, 192, inverse=True)), d, [h, w], True)[0]
    hyp[:, :, :, w // 2:] = cuda(425 + 500 * torch.rand(1, d, h, w - w // 2, generator=g))  # half staged, half direct
    half = [ops.features_nhwc(feats[0])] + [ops.features_nhwc_f16(f) for f in feats[1:]]
    full = ops.warp_corr(half, rt, hyp, layout="h16")
    for r0, r1 in ((0, 24), (24, 72), (40, h)):
        band = ops.warp_corr([half[0][:, :, r0:r1]] + half[1:], rt, hyp[:, :, r0:r1].contiguous(), layout="h16", row0=r0)
        assert torch.equal(band, full[:, :, :, r0:r1])


def _sharded_worker(rank, world, port, out_dir):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dmvsnet_b200 import MVSNet, parallel, synthetic as syn
        H, W, n, nd, ratios = 512, 320, 5, [16, 8, 8], [4, 2, 1]
        net = MVSNet(nd, ratios, inverse_depth=True)
        net.load_state_dict(syn.ridge_regnet_state(net.state_dict(), seed=3))
        net = net.to("cuda:%d" % rank).eval()
        proj = syn.make_proj_matrices(H, W, n, 1, num_stages=3)
        imgs = syn.make_scene_images(H, W, n, proj["stage3"], seed=3).to("cuda:%d" % rank)
        dv = syn.make_depth_values(1, 192, inverse=True).to("cuda:%d" % rank)
        with torch.no_grad():
            want = net(imgs, proj, dv)
            got = parallel.infer_view_sharded(net, imgs, proj, dv)
        ok = all(torch.equal(got["stage%d" % s][k], want["stage%d" % s][k]) for s in (1, 2, 3)
                 for k in ("depth", "photometric_confidence", "photometric_confidence_refine", "depth_values_c", "depth_sub_plus"))
        open(os.path.join(out_dir, "rank%d" % rank), "w").write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_single_view_sharded_mode_is_bit_identical_on_two_gpus(tmp_path):
    """FeatureNet by view + fp16 source all-gather + row-band cascade over NCCL == the one-GPU forward, every re-assembled map of
    every stage bit for bit (8 ranks at T&T size: tools/check_view_sharded_nccl.py, profiles/r2g_view_sharded_n8.txt)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(str(tmp_path / ("rank%d" % r))).read() == "ok"
