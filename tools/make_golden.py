#!/usr/bin/env python
"""Generate tests/golden/*.npz from the LIVE reference and pin the oracle against it.

Runs only in the build container (needs /root/reference; it is absent on the GPU box).
For every case below it
  1. regenerates the synthetic inputs from ``dmvsnet_b200.synthetic`` (seeded),
  2. runs the imported, unmodified reference ``networks.mvsnet.MVSNet`` on PyTorch-CPU fp32,
     capturing the seams (cost volume, logits, heads) with forward hooks,
  3. runs ``oracle/dmvs_oracle.py`` on the same inputs and prints / asserts the differences,
  4. stores the REFERENCE outputs (not the oracle's) as the fixture.

Usage:  PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF = os.environ.get("DMVS_REFERENCE", "/root/reference")

from dmvsnet_b200 import ops, synthetic as syn  # noqa: E402
from oracle import dmvs_oracle as O  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES  # noqa: E402  (shared with the tests)

SEAM_KEYS = ["depth_values", "depth_sub_plus", "depth_values_c", "photometric_confidence", "depth",
             "photometric_confidence_refine", "depth_sub_plus_refine", "interval"]


def import_reference():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        from networks import mvsnet as MV  # type: ignore
        from networks import module as MM  # type: ignore
    return MV, MM


def build_reference(MV, case, seed=0):
    with contextlib.redirect_stdout(io.StringIO()):
        net = MV.MVSNet(ndepths=case["ndepths"], depth_interval_ratio=case["ratios"], inverse_depth=case["inverse"])
    state = syn.randomise_regnet_state(net.state_dict(), seed=seed)
    net.load_state_dict(state)
    net.eval()
    return net, state


def case_inputs(case, seed=0):
    H, W, N, B = case["H"], case["W"], case["views"], case["batch"]
    ns = len(case["ndepths"])
    proj = syn.make_proj_matrices(H, W, N, B, num_stages=ns)
    dv = syn.make_depth_values(B, 192, inverse=case["inverse"])
    if case["mode"] == "images":
        return dict(imgs=syn.make_images(H, W, N, B, seed=seed), proj=proj, depth_values=dv)
    feats = syn.make_stage_features(H, W, N, B, seed=seed, num_stages=ns)
    return dict(features=feats, imgs=torch.zeros(B, N, 3, H, W), proj=proj, depth_values=dv)


class _Replay(torch.nn.Module):
    """Stands in for FeatureNet: returns the precomputed per-view feature dicts in call order."""

    def __init__(self, feats):
        super().__init__()
        self._it = iter(feats)

    def forward(self, img):
        return next(self._it)


def run_reference(net, inp, case):
    seams = {}
    hooks = []

    def grab(name):
        store = []
        seams[name] = store
        return lambda mod, args, out: store.append(out.detach().clone())

    hooks.append(net.cost_aggregation.register_forward_hook(grab("cost")))
    for i, m in enumerate(net.cost_regularization):
        hooks.append(m.register_forward_hook(grab("logits%d" % i)))
    for i, m in enumerate(net.cost_regularization_refine):
        hooks.append(m.register_forward_hook(grab("logits_c%d" % i)))
    orig_feature = net.feature
    if "features" in inp:
        net.feature = _Replay(inp["features"])  # inject precomputed per-view feature dicts
    with torch.no_grad():
        out = net(inp["imgs"], inp["proj"], inp["depth_values"])
    net.feature = orig_feature
    for h in hooks:
        h.remove()
    flat = {}
    for s in range(len(case["ndepths"])):
        st = out["stage%d" % (s + 1)]
        for k in SEAM_KEYS:
            flat["s%d_%s" % (s + 1, k)] = st[k]
        flat["s%d_cost" % (s + 1)] = seams["cost"][2 * s]
        flat["s%d_cost_c" % (s + 1)] = seams["cost"][2 * s + 1]
        flat["s%d_logits" % (s + 1)] = seams["logits%d" % s][0]
        flat["s%d_logits_c" % (s + 1)] = seams["logits_c%d" % s][0]
        # the homographies exactly as the reference computed them on this host (bit-equal, see test_relative_projections_match_oracle)
        flat["s%d_rt" % (s + 1)] = ops.relative_projections(inp["proj"]["stage%d" % (s + 1)])
    return flat


def run_oracle(state, inp, case):
    with torch.no_grad():
        if "features" in inp:
            out = O.cascade_forward(inp["features"], inp["proj"], inp["depth_values"], state, case["ndepths"],
                                    case["ratios"], case["inverse"], (case["H"], case["W"]), keep_seams=True)
        else:
            out = O.mvsnet_forward(inp["imgs"], inp["proj"], inp["depth_values"], state, case["ndepths"],
                                   case["ratios"], case["inverse"], keep_seams=True)
    flat = {}
    for s in range(len(case["ndepths"])):
        st = out["stage%d" % (s + 1)]
        for k in SEAM_KEYS:
            flat["s%d_%s" % (s + 1, k)] = st[k]
        for k in ("cost", "cost_c", "logits", "logits_c"):
            flat["s%d_%s" % (s + 1, k)] = st["_" + k]
    return flat


def relerr(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def edge_case_warp(MV):
    """W1-only fixture: rotated rig, out-of-frustum samples, z<0 and exact z==0 hypotheses."""
    g = torch.Generator().manual_seed(7)
    B, C, h, w, D, N = 2, 8, 24, 40, 6, 3
    feats = [torch.randn(B, C, h, w, generator=g) for _ in range(N)]
    proj = syn.make_proj_matrices(h * 4, w * 4, N, B, num_stages=1)["stage1"]
    hyp = 425.0 + 600.0 * torch.rand(B, D, h, w, generator=g)
    hyp[:, 0, :4, :] = -50.0          # behind-camera samples are NOT masked by the reference
    hyp[:, 1, 4:6, :] = 30.0          # far out of frustum -> zero padding
    # exact Z == 0 for source 1 at pixel (10, 20): solve r2.(x,y,1) d + t2 = 0 in float32
    with torch.no_grad():
        rp = O.compose_projection(proj[:, 0])
        rot, trans = O.relative_projection(O.compose_projection(proj[:, 1]), rp)
        r2 = rot[0, 2, 0] * 20.0 + rot[0, 2, 1] * 10.0 + rot[0, 2, 2]
        hyp[0, 2, 10, 20] = -trans[0, 2] / r2
    agg = MV.CostAgg("variance", None).eval()
    with torch.no_grad():
        cost = agg(feats, proj, hyp, 0)
        ocost = O.warp_corr(feats, proj, hyp)
    print("  warp_edge: oracle vs reference cost rel-Linf %.3e  (nan count ref %d oracle %d)" % (
        relerr(torch.nan_to_num(ocost), torch.nan_to_num(cost)), int(torch.isnan(cost).sum()), int(torch.isnan(ocost).sum())))
    return dict(cost=cost.numpy(), hyp=hyp.numpy(), proj=proj.numpy(), **{"feat%d" % i: f.numpy() for i, f in enumerate(feats)})


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    MV, MM = import_reference()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    worst = {}
    for name, case in CASES.items():
        net, state = build_reference(MV, case)
        inp = case_inputs(case)
        ref = run_reference(net, inp, case)
        ora = run_oracle(state, inp, case)
        print("case %s" % name)
        for k in sorted(ref):
            if k.endswith("_rt"):
                continue
            e = relerr(ora[k], ref[k])
            worst[k.split("_", 1)[1]] = max(worst.get(k.split("_", 1)[1], 0.0), e)
            print("  %-34s shape %-22s ref|max| %.4g  oracle rel-Linf %.3e" % (k, tuple(ref[k].shape), float(ref[k].abs().max()), e))
        np.savez(os.path.join(outdir, name + ".npz"), **{k: v.numpy().astype(np.float32) for k, v in ref.items()})
        # softmax peakedness: confirms the weights are non-degenerate
        for s in range(len(case["ndepths"])):
            p = torch.softmax(ref["s%d_logits" % (s + 1)], 2).max(2)[0].mean()
            hc = ref["s%d_depth_values_c" % (s + 1)]
            print("  stage%d mean max-prob %.3f (uniform %.3f); refine hypotheses in [%.1f, %.1f], final depth in [%.1f, %.1f]" % (
                s + 1, float(p), 1.0 / case["ndepths"][s], float(hc.min()), float(hc.max()),
                float(ref["s%d_depth" % (s + 1)].min()), float(ref["s%d_depth" % (s + 1)].max())))
    np.savez(os.path.join(outdir, "warp_edge.npz"), **edge_case_warp(MV))
    print("worst oracle-vs-reference rel-Linf per seam:")
    for k, v in sorted(worst.items()):
        print("  %-32s %.3e" % (k, v))
    bit_exact = ["depth_values"]
    for k in bit_exact:
        assert worst[k] == 0.0, k
    assert all(v < 2e-4 for v in worst.values()), worst


if __name__ == "__main__":
    main()
