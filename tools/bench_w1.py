"""W1 micro-benchmark (run on the GPU box): both kernels (channel-last gather / reference-layout gather) on the six
DTU launches of one cascade step, with hypotheses of two kinds:

  smooth  the previous stage's depth is a slanted plane (+ a step edge): what a trained network hands down
  noise   the previous stage's depth is white noise over the whole depth range: what randomly initialised
          regularisation nets hand down (bench.py's workload)

Prints per launch: ms, algorithmic GB/s, fraction of the measured HBM peak.  Timing: CUDA events around 10 launches
after 3 warm-ups; the feature maps of one launch set (>= 150 MB) exceed nothing in L2 terms for stage 1 - so a
256 MB scratch write flushes L2 between launches when --flush is given.

    python tools/bench_w1.py [--config dtu] [--layouts nhwc,nchw] [--kinds smooth,noise] [--flush]

The pseudo-layout "bwd" times the backward kernel (dmvs_warp_corr_backward_f32, gradients to all feature maps) on the
same launches; "--config train" is the reference's 640x512 training crop.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dmvsnet_b200 import ops, synthetic as syn  # noqa: E402

CONFIGS = {"dtu": (1184, 1600, 5, [48, 32, 8]), "bmvs": (576, 768, 7, [48, 32, 8]), "tnt": (1056, 1920, 11, [48, 32, 8]),
           "small": (256, 320, 5, [48, 32, 8]), "train": (512, 640, 5, [48, 32, 8])}  # train: the reference's DTU training crop


def depth_map(kind, h, w, g, dev):
    if kind == "noise":
        return (425 + 500 * torch.rand(1, h, w, generator=g)).to(dev)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    d = 560 + 160 * xs / w + 80 * ys / h
    d = d + 60.0 * ((xs / w + 0.3 * ys / h) > 0.6)  # one depth discontinuity
    return d[None].to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="dtu")
    ap.add_argument("--layouts", default="h16,staged,nhwc,nchw")
    ap.add_argument("--kinds", default="smooth,noise")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--once", action="store_true", help="one launch per case between cudaProfilerStart/Stop (for ncu)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    H, W, views, nd = CONFIGS[args.config]
    ratios = [4, 2, 1]
    peak = 6538.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    g = torch.Generator().manual_seed(0)
    proj = syn.make_proj_matrices(H, W, views, 1, num_stages=3)
    dv = syn.make_depth_values(1, 192, inverse=True).to(dev)
    interval = (dv[0, -1] - dv[0, 0]) / dv.size(1)
    scratch = torch.empty(64 * 1024 * 1024, device=dev) if args.flush else None
    rows = []
    if args.once:
        torch.cuda.cudart().cudaProfilerStart()
    for s in range(3):
        scale = 2 ** (2 - s)
        h, w, c = H // scale, W // scale, 32 >> s
        rt = ops.relative_projections(proj["stage%d" % (s + 1)]).to(dev)
        feats = [torch.randn(1, c, h, w, generator=g).to(dev) for _ in range(views)]
        feats_cl = [feats[0]] + [ops.features_nhwc(f) for f in feats[1:]]
        feats_h16 = [feats_cl[0]] + [ops.features_nhwc_f16(f) for f in feats_cl[1:]]
        for kind in args.kinds.split(","):
            if s == 0:
                hyp, iv = ops.hypotheses_first(dv, nd[0], [h, w], True)
                last = depth_map(kind, h, w, g, dev)
            else:
                last = depth_map(kind, h // 2, w // 2, g, dev)
                hyp, iv = ops.hypotheses_next(last, nd[s], ratios[s] * interval, [h, w], True)
                last = depth_map(kind, h, w, g, dev)
            # refine hypotheses: 4 values around the regressed depth (mvsnet.py:33-56 picks them from a 6-stack)
            step = float(iv) if kind == "smooth" else 40.0
            hyp_c = torch.stack([last + step * (k - 1.5) for k in range(4)], 1).contiguous()
            for name, hy in (("main", hyp), ("refine", hyp_c)):
                d = hy.shape[1]
                for layout in args.layouts.split(","):
                    alg = 4 * h * w * (views * c + 3 * d)
                    fs = feats if layout == "nchw" else (feats_h16 if layout == "h16" else feats_cl)
                    run = lambda: ops.warp_corr(fs, rt, hy, layout=layout)  # noqa: E731
                    if layout == "bwd":
                        fs_all = [ops.features_nhwc(feats[0])] + feats_cl[1:]
                        gout = torch.randn(1, 2, d, h, w, generator=g).to(dev)
                        alg = 4 * h * w * (2 * views * c + 3 * d)
                        run = lambda: ops.warp_corr_backward(fs_all, rt, hy, gout)  # noqa: E731
                    if args.once:
                        run()
                        continue
                    for _ in range(3):
                        out = run()
                    ms = 0.0
                    for _ in range(args.iters):
                        if scratch is not None:
                            scratch.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        out = run()
                        e1.record()
                        torch.cuda.synchronize()
                        ms += e0.elapsed_time(e1)
                    ms /= args.iters
                    rows.append((s + 1, name, kind, layout, c, d, h, w, ms, alg / ms / 1e6))
                    extra = ""
                    if layout == "staged":
                        extra = "  fallback tiles %.1f %%" % (100.0 * float(ops.LAST_W1_FLAGS.float().mean()))
                    print("stage%d %-6s %-6s %-6s C=%-2d D=%-2d %4dx%-4d  %7.3f ms  %7.1f GB/s  %.3f of HBM peak%s" %
                          (s + 1, name, kind, layout, c, d, h, w, ms, alg / ms / 1e6, alg / ms / 1e6 / peak, extra), flush=True)
                    del out
        # transposition cost of this stage's source maps (both feature sets)
        if not args.once:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                for f in feats[1:]:
                    ops.features_nhwc(f)
            e1.record()
            torch.cuda.synchronize()
            print("stage%d repack of %d source maps: %.3f ms" % (s + 1, views - 1, e0.elapsed_time(e1) / 5), flush=True)
    if args.once:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    for kind in args.kinds.split(","):
        for layout in args.layouts.split(","):
            sel = [r for r in rows if r[2] == kind and r[3] == layout]
            ms = sum(r[8] for r in sel)
            alg = sum(r[9] * r[8] * 1e6 for r in sel)
            print("TOTAL %-6s %-4s  %7.3f ms per step  %7.1f GB/s pooled  %.3f of HBM peak" % (kind, layout, ms, alg / ms / 1e6, alg / ms / 1e6 / peak))


if __name__ == "__main__":
    main()
