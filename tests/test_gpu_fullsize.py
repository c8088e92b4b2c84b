"""Full-size parity: the CUDA path against the oracle at the stage grids of BASELINE.json's configurations.

The small-shape tests (test_gpu_parity.py) cover the edge cases; these cover what only shows at size: the persistent tile
schedulers, TMA boxes and 64-bit offsets of the tensor convs at 592x800x32 / 1184x1600x8, the staged W1 bounding-box logic at
296x400x48, source loops of N = 7 / 11 views (configs 3 / 4) and the D = 64 / 16 head instantiations (config 5).  The oracle
runs on the host cores of the GPU box (seconds per case; the whole file a few minutes).

Shapes: /root/reference/scripts/dtu_test.sh:8-30 (1600x1184, N=5, 48/32/8), scripts/tank_test.sh:8-24 (1920x1056, N=11),
BASELINE.json configs 3 (768x576, N=7) and 5 (D = 64/32/16).
"""
import pytest
import torch

from conftest import rel_linf
from oracle import dmvs_oracle as O
from test_oracle_golden import _c_warp_corr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _lib(native_lib):
    return native_lib


def cuda(t):
    return t.to(DEV)


def _mixed_depth(h, w, g):
    """Left part: slanted plane with a step edge (what a trained network hands down); right third: white noise over the
    whole depth range (what random regularisation weights hand down) - both gather regimes in one map."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    d = 560 + 160 * xs / w + 80 * ys / h + 60.0 * ((xs / w + 0.3 * ys / h) > 0.45)
    noise = 425 + 500 * torch.rand(h, w, generator=g)
    return torch.where(xs > 0.66 * w, noise, d)[None]


W1_FULL = [  # (config, stage, H, W, views, C, D, refine)
    ("dtu", 1, 1184, 1600, 5, 32, 48, False), ("dtu", 2, 1184, 1600, 5, 16, 32, False), ("dtu", 3, 1184, 1600, 5, 8, 8, False),
    ("dtu", 3, 1184, 1600, 5, 8, 4, True), ("tnt", 3, 1056, 1920, 11, 8, 4, True), ("tnt", 1, 1056, 1920, 11, 32, 48, False),
    ("bmvs", 2, 576, 768, 7, 16, 32, False), ("bmvs", 1, 576, 768, 7, 32, 4, True),
]


@pytest.mark.parametrize("cfg,stage,H,W,views,c,d,refine", W1_FULL)
def test_warp_corr_full_size_vs_c_oracle(cfg, stage, H, W, views, c, d, refine, c_oracle):
    """Every W1 kernel at a full stage grid against the plain-C restatement (oracle/warp_corr_ref.c, OpenMP)."""
    from dmvsnet_b200 import ops, synthetic as syn
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
    want = _c_warp_corr(c_oracle, feats, rt, hyp)
    dfeats, drt, dhyp = [cuda(f) for f in feats], cuda(rt), cuda(hyp)
    for layout in ("nhwc", "staged", "nchw"):
        got = ops.warp_corr(dfeats, drt, dhyp, layout=layout)
        err = rel_linf(got, want)
        assert err < 2e-6, (cfg, stage, layout, err)
        del got


@pytest.mark.parametrize("refine,d,h,w", [(False, 32, 592, 800), (True, 4, 592, 800), (False, 8, 1184, 1600), (False, 48, 264, 480),
                                          (True, 4, 1056, 1920)])
def test_regnet_full_size_vs_oracle(refine, d, h, w):
    """One CostRegNet / CostRegNet_refine (both U-Net branches) at a full stage grid against the oracle (torch-CPU conv3d):
    DTU stage 2 main + refine (module.py:358-436), DTU stage 3, T&T stage 1 and the T&T stage-3 refine net."""
    from dmvsnet_b200 import MVSNet, synthetic as syn
    net = MVSNet([8, 8, 8], [4, 2, 1])
    state = syn.randomise_regnet_state(net.state_dict(), seed=2)
    net.load_state_dict(state)
    net = net.to(DEV).eval()
    mod = (net.cost_regularization_refine if refine else net.cost_regularization)[1]
    prefix = "cost_regularization%s.1." % ("_refine" if refine else "")
    g = torch.Generator().manual_seed(d + h)
    x = torch.randn(1, 2, d, h, w, generator=g)
    with torch.no_grad():
        want = O.regnet(x, O._sub(state, prefix), refine=refine)
        got = mod(cuda(x)).cpu()
    err = rel_linf(got, want)
    assert err < 1e-4, err


@pytest.mark.parametrize("d,h,w", [(64, 96, 128), (16, 384, 512), (32, 592, 800)])
def test_depth_head_config5_depths_vs_oracle(d, h, w):
    """depth_head_kernel<D> instantiations of BASELINE config 5 (D = 64 / 16) and the DTU stage-2 grid."""
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(d)
    logits = 6 * torch.randn(1, 4, d, h, w, generator=g)
    hyp = 425 + 500 * torch.rand(1, d, h, w, generator=g).sort(1)[0]
    interval = torch.tensor(10.77)
    want = O.depth_head(logits, hyp, interval)
    prob, d4, hyp_c, conf = ops.depth_head(cuda(logits), cuda(hyp), cuda(interval))
    assert rel_linf(prob, want["prob_volume"]) < 1e-5
    assert rel_linf(d4, want["depth_sub_plus"]) < 1e-6
    assert rel_linf(hyp_c, want["depth_values_c"]) < 1e-5
    assert float((conf.cpu() - want["photometric_confidence"]).abs().max()) < 1e-4


def _seam_err(got, ref, how):
    diff = (got.cpu() - ref).abs()
    e = {"rel": diff / ref.abs().clamp_min(1.0), "lin": diff / ref.abs().max(), "abs": diff}[how].flatten()
    return float(e.max()), float(e.kthvalue(max(1, int(0.999 * e.numel())))[0]), float(e.mean())


@pytest.mark.parametrize("cfg,H,W,views,nd", [("dtu", 1184, 1600, 5, [48, 32, 8]), ("bmvs", 576, 768, 7, [48, 32, 8])])
def test_cascade_full_size_vs_oracle(cfg, H, W, views, nd):
    """BASELINE config 2 (the benchmarked configuration) and config 3 against ``O.cascade_forward`` on the same host
    (mvsnet.py:188-260), all 11 seams x 3 stages with the tolerances of the small-shape cascade test.

    Seam by seam with the ORACLE's value as the input of every step (hypotheses from the oracle's previous depth, cost
    volume from the oracle's hypotheses, heads from the oracle's logits ...): with the App. D random weights the cascade is a
    chaotic map at this size - a peaked softmax at a random plane per pixel, min/max selections and the 3a-2b extrapolation
    (mvsnet.py:42-45) turn a 1e-5 difference into another plane at a handful of the 1.9 M pixels, and every later seam inherits
    it - so a free-running comparison measures the conditioning of the random network, not the kernels.  The free-running
    cascade is still run and its errors printed (max / p99.9 / mean); its well-conditioned stage-1 main seams are asserted.
    ``test_cascade_full_size_conditioned_weights`` is the free-running check on a network that behaves like a trained one."""
    from dmvsnet_b200 import MVSNet, ops, synthetic as syn
    from test_gpu_parity import CASCADE_SEAMS
    tol = {k: (t, how) for k, t, how in CASCADE_SEAMS}
    ratios = [4, 2, 1]
    net = MVSNet(nd, ratios, inverse_depth=True)
    state = syn.randomise_regnet_state(net.state_dict(), seed=1)
    net.load_state_dict(state)
    net = net.to(DEV).eval()
    net.w1_precision = "fp32"
    feats = syn.make_stage_features(H, W, views, 1, seed=3)
    proj = syn.make_proj_matrices(H, W, views, 1, num_stages=3)
    dv = syn.make_depth_values(1, 192, inverse=True)
    dfeats = [{k: cuda(v) for k, v in f.items()} for f in feats]
    with torch.no_grad():
        want = O.cascade_forward(feats, proj, dv, state, nd, ratios, True, (H, W), keep_seams=True)
        free = net.cascade(dfeats, proj, cuda(dv), (H, W), keep_seams=True)
    report, bad = [], []

    def check(stage, seam, got, ref=None, assert_it=True, tag=""):
        ref = ref if ref is not None else (want["stage%d" % stage]["_" + seam] if "_" + seam in want["stage%d" % stage] else want["stage%d" % stage][seam])
        t, how = tol[seam]
        if seam == "depth_values" and stage > 1:
            t = 5e-5  # inverse-depth sampling around a white-noise previous depth: 1 / (1/lo + k * step) cancels at a few pixels
        mx, p999, mean = _seam_err(got, ref, how)
        report.append("stage%d %-30s %s %s max %.2e  p99.9 %.2e  mean %.2e (tol %.0e)" % (stage, seam, tag, how, mx, p999, mean, t))
        if assert_it and not mx < t:
            bad.append(report[-1])

    depth_interval = (dv[0, -1] - dv[0, 0]) / dv.size(1)
    with torch.no_grad():
        for s in range(3):
            stage, name = s + 1, "stage%d" % (s + 1)
            ws = want[name]
            h, w = H >> (2 - s), W >> (2 - s)
            rt = cuda(ops.relative_projections(proj[name]))
            # S1 from the oracle's previous depth
            if s == 0:
                hyp, interval = ops.hypotheses_first(cuda(dv), nd[s], [h, w], True)
            else:
                hyp, interval = ops.hypotheses_next(cuda(want["stage%d" % s]["depth"]), nd[s], cuda(ratios[s] * depth_interval), [h, w], True)
            check(stage, "depth_values", hyp, tag="forced")
            # W1 (fp32 volume + the conv0 cells the cascade really feeds) from the oracle's hypotheses; R1 from those cells
            o_hyp = cuda(ws["depth_values"])
            cost, cells = net.cost_aggregation.forward_fused([f[name] for f in dfeats], o_hyp, rt, want_f32=True, coherent=(s == 0))
            check(stage, "cost", cost, tag="forced")
            logits = net.cost_regularization[s](None, cost_cells=cells)
            check(stage, "logits", logits, tag="forced")
            del cost, cells, logits
            # E1 from the oracle's logits
            prob, d4, hyp_c, conf = ops.depth_head(cuda(ws["_logits"]), o_hyp, cuda(ws["interval"]))
            check(stage, "depth_sub_plus", d4, tag="forced")
            check(stage, "depth_values_c", hyp_c, tag="forced")
            check(stage, "photometric_confidence", conf, tag="forced")
            assert rel_linf(prob, ws["prob_volume"]) < 1e-5
            del prob
            # refine pass from the oracle's refine hypotheses
            o_hyp_c = cuda(ws["depth_values_c"])
            cost_c, cells_c = net.cost_aggregation.forward_fused([f[name + "_c"] for f in dfeats], o_hyp_c, rt, want_f32=True)
            check(stage, "cost_c", cost_c, tag="forced")
            logits_c = net.cost_regularization_refine[s](None, cost_cells=cells_c)
            check(stage, "logits_c", logits_c, tag="forced")
            del cost_c, cells_c, logits_c
            depth, conf_r, d4r = ops.refine_head(cuda(ws["_logits_c"]), o_hyp_c, cuda(ws["interval"]), 5.0)
            check(stage, "depth_sub_plus_refine", d4r, tag="forced")
            check(stage, "depth", depth, tag="forced")
            check(stage, "photometric_confidence_refine", conf_r, tag="forced")
    # the free-running cascade: report everything, assert what is well conditioned
    for s in range(3):
        st = free["stage%d" % (s + 1)]
        for seam, _, _ in CASCADE_SEAMS:
            got = st["_" + seam] if "_" + seam in st else st[seam]
            check(s + 1, seam, got, assert_it=(s == 0 and seam in ("depth_values", "cost", "logits", "depth_sub_plus")), tag="free  ")
    print("\n".join(report))
    assert not bad, "\n".join(["seams out of tolerance:"] + bad + ["all seams:"] + report)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("cfg,H,W,views,nd", [("dtu", 1184, 1600, 5, [48, 32, 8])])
def test_cascade_full_size_conditioned_weights(cfg, H, W, views, nd, precision):
    """FREE-RUNNING cascade at BASELINE config 2 on the workload bench.py times: photo-consistent feature maps of a tilted
    plane (synthetic.make_scene_features) and regularisation nets that follow the cost ridge like trained ones
    (synthetic.ridge_regnet_state).  Every stage's final depth within the 1e-3 contract of the oracle's, for the exact W1
    kernels and for the fp16-staged one, and close to the depth of the rendered scene."""
    from dmvsnet_b200 import MVSNet, synthetic as syn
    ratios = [4, 2, 1]
    net = MVSNet(nd, ratios, inverse_depth=True)
    state = syn.ridge_regnet_state(net.state_dict(), seed=1)
    net.load_state_dict(state)
    net = net.to(DEV).eval()
    net.w1_precision = precision
    proj = syn.make_proj_matrices(H, W, views, 1, num_stages=3)
    feats = syn.make_scene_features(H, W, views, proj, seed=3)
    dv = syn.make_depth_values(1, 192, inverse=True)
    with torch.no_grad():
        want = O.cascade_forward(feats, proj, dv, state, nd, ratios, True, (H, W))
        out = net.cascade([{k: cuda(v) for k, v in f.items()} for f in feats], proj, cuda(dv), (H, W))
    scene = syn.scene_depth(H, W, proj["stage3"])
    report, bad = [], []
    for s in range(3):
        name = "stage%d" % (s + 1)
        for seam in ("depth_sub_plus", "depth_values_c", "depth"):
            got, ref = out[name][seam].cpu(), want[name][seam]
            e = ((got - ref).abs() / ref.abs().clamp_min(1.0)).flatten()
            p999 = float(e.kthvalue(max(1, int(0.999 * e.numel())))[0])
            beyond = float((e > 1e-3).float().mean())
            report.append("%s %-16s max %.2e  p99.9 %.2e  mean %.2e  pixels beyond 1e-3: %.1e" % (name, seam, float(e.max()), p999, float(e.mean()), beyond))
            # the contract is 1e-3 relative on the regressed depth.  Over 1.9 M pixels a handful sit on a near-tie of two softmax
            # peaks (image borders, where source views drop out of the frustum) and move by more whatever the arithmetic - the
            # oracle on another host does the same - so: 99.9 % of the pixels ten times inside the contract, < 1e-4 of them outside
            if not (p999 < 2e-4 and beyond < 1e-4):
                bad.append(report[-1])
    off = float(((out["depth"].cpu()[0] - scene).abs() / scene).mean())
    report.append("final depth vs the rendered plane: mean rel %.2e" % off)
    print("\n".join(report))
    assert not bad and off < 5e-3, "\n".join(report)
