"""Deterministic synthetic inputs for the cost-volume hot path.

No dataset or checkpoint is available offline, so every test, fixture and
bench line is driven from here.  The generator follows SURVEY.md §8(d):

* DTU-like pinhole intrinsics scaled to the requested image size, mirroring
  the ``/4`` at reference ``datasets/general_eval.py:69`` and the x2 / x4 per
  stage at ``datasets/general_eval.py:190-192``;
* a look-at camera rig (source cameras 100-300 mm off the reference, optical
  axes through (0, 0, 680)) so that ~90 % of plane-sweep samples land inside
  the source images;
* ``depth_values`` in the two forms the reference datasets emit
  (``datasets/dtu_yao.py:68,167`` linear, ``datasets/general_eval.py:178-181``
  inverse);
* a non-degenerate weight recipe: default-initialised
  weights give an exactly uniform softmax, which would let any warp kernel pass
  a final-depth parity check.

Everything is generated on the CPU with a seeded ``torch.Generator`` so the
GPU box (same image, same torch build) regenerates identical tensors and the
committed golden outputs under ``tests/golden/`` stay valid.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch

# camera centres of the source views in the reference frame, millimetres
_RIG_CENTRES = [
    (100.0, 0.0, 0.0), (-100.0, 0.0, 0.0), (0.0, 100.0, 0.0), (0.0, -100.0, 0.0),
    (200.0, 0.0, 0.0), (-200.0, 0.0, 0.0), (-200.0, 50.0, 0.0), (200.0, -50.0, 0.0),
    (150.0, 150.0, 0.0), (-150.0, -150.0, 0.0), (300.0, 0.0, 0.0), (-300.0, 0.0, 0.0),
]
_LOOK_AT = (0.0, 0.0, 680.0)

FEATURE_CHANNELS = (32, 16, 8)  # stage1..3, reference networks/module.py:301-311


def _normalise(v: torch.Tensor) -> torch.Tensor:
    return v / v.norm()


def look_at_extrinsic(centre: Sequence[float]) -> torch.Tensor:
    """World(=reference camera)-to-camera 4x4 for a camera at ``centre`` looking at _LOOK_AT."""
    c = torch.tensor(centre, dtype=torch.float64)
    z = _normalise(torch.tensor(_LOOK_AT, dtype=torch.float64) - c)
    x = _normalise(torch.linalg.cross(torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64), z))
    y = torch.linalg.cross(z, x)
    rot = torch.stack([x, y, z])
    ext = torch.eye(4, dtype=torch.float64)
    ext[:3, :3] = rot
    ext[:3, 3] = -rot @ c
    return ext.float()


def stage1_intrinsics(height: int, width: int) -> torch.Tensor:
    k = torch.eye(3, dtype=torch.float32)
    k[0, 0] = k[1, 1] = 2892.33 * (width / 1600.0) / 4.0
    k[0, 2] = 823.2 * (width / 1600.0) / 4.0
    k[1, 2] = 619.07 * (height / 1200.0) / 4.0
    return k


def make_proj_matrices(height: int, width: int, num_views: int, batch: int = 1,
                       num_stages: int = 3) -> Dict[str, torch.Tensor]:
    """``{"stageK": [B, N, 2, 4, 4]}``; [:, :, 0] extrinsic, [:, :, 1, :3, :3] intrinsic."""
    assert num_views - 1 <= len(_RIG_CENTRES), "rig has %d source poses" % len(_RIG_CENTRES)
    k1 = stage1_intrinsics(height, width)
    per_view = []
    for v in range(num_views):
        pm = torch.zeros(2, 4, 4)
        pm[0] = torch.eye(4) if v == 0 else look_at_extrinsic(_RIG_CENTRES[v - 1])
        pm[1, :3, :3] = k1
        per_view.append(pm)
    base = torch.stack(per_view)  # [N,2,4,4]
    out = {}
    for s in range(num_stages):
        pm = base.clone()
        pm[:, 1, :2, :] = base[:, 1, :2, :] * float(2 ** s)
        pm = pm.unsqueeze(0).repeat(batch, 1, 1, 1, 1).contiguous()
        if batch > 1:
            # give later batch entries a slightly different rig so that batching bugs show
            for b in range(1, batch):
                pm[b, 1:, 0, :3, 3] *= (1.0 + 0.05 * b)
        out["stage%d" % (s + 1)] = pm
    return out


def make_depth_values(batch: int = 1, numdepth: int = 192, inverse: bool = False,
                      depth_min: float = 425.0, depth_interval: float = 2.5 * 1.06) -> torch.Tensor:
    if inverse:
        depth_end = depth_interval * numdepth + depth_min
        inv = torch.linspace(1.0 / depth_min, 1.0 / depth_end, numdepth + 1, dtype=torch.float64)[:-1]
        dv = (1.0 / inv).float()
    else:
        dv = (depth_min + depth_interval * torch.arange(numdepth, dtype=torch.float64)).float()
    return dv.unsqueeze(0).repeat(batch, 1).contiguous()


def make_images(height: int, width: int, num_views: int, batch: int = 1, seed: int = 0, natural: bool = False) -> torch.Tensor:
    """``natural=False``: iid U[0,1) pixels (the parity fixtures).  ``natural=True``: a 1/f-like multi-octave texture shared by
    all views plus 2 % pixel noise - photographs are spatially smooth, and so are the depth maps (hence the gather
    locality of the stage-2/3 warps) that a network regresses from them; white noise would misrepresent both."""
    g = torch.Generator().manual_seed(seed)
    if not natural:
        return torch.rand(batch, num_views, 3, height, width, generator=g)
    tex = torch.zeros(batch, 3, height, width)
    amp, octave = 1.0, 0
    while (height >> octave) >= 4 and octave < 8:
        hh, ww = max(2, height >> octave), max(2, width >> octave)
        layer = torch.randn(batch, 3, hh, ww, generator=g)
        tex = tex + amp * torch.nn.functional.interpolate(layer, size=(height, width), mode="bilinear", align_corners=False)
        amp, octave = amp * 1.6, octave + 1
    tex = (tex - tex.mean()) / (4.0 * tex.std()) + 0.5
    views = [tex + 0.02 * torch.randn(batch, 3, height, width, generator=g) for _ in range(num_views)]
    return torch.stack(views, 1).clamp_(0.0, 1.0).contiguous()


def _fractal_texture(height: int, width: int, g: torch.Generator, channels: int = 3) -> torch.Tensor:
    """1/f-like multi-octave texture [channels, height, width], roughly zero mean / unit variance."""
    tex = torch.zeros(1, channels, height, width)
    amp, octave = 1.0, 0
    while (min(height, width) >> octave) >= 4 and octave < 9:
        hh, ww = max(2, height >> octave), max(2, width >> octave)
        layer = torch.randn(1, channels, hh, ww, generator=g)
        tex = tex + amp * torch.nn.functional.interpolate(layer, size=(height, width), mode="bilinear", align_corners=False)
        amp, octave = amp * 1.5, octave + 1
    return ((tex - tex.mean()) / tex.std())[0]


def scene_depth(height: int, width: int, proj_full: torch.Tensor) -> torch.Tensor:
    """Depth map [height, width] (reference view, millimetres) of the surface ``make_scene_images`` renders: the plane
    n . X = c through (0, 0, 640) tilted about both axes, in the reference camera frame (extrinsic = identity)."""
    k = proj_full[0, 0, 1, :3, :3].double()
    n, c = _scene_plane()
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float64), torch.arange(width, dtype=torch.float64), indexing="ij")
    rays = torch.linalg.solve(k, torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(3, -1))
    lam = c / (n[:, None] * rays).sum(0)
    return (lam * rays[2]).reshape(height, width).float()


def _scene_plane():
    n = torch.tensor([0.22, -0.12, 1.0], dtype=torch.float64)
    n = n / n.norm()
    return n, float(n[2] * 640.0)


def scene_depth_view(height: int, width: int, proj_full: torch.Tensor, view: int) -> torch.Tensor:
    """Depth map [height, width] of the rendered plane as camera ``view`` of the rig sees it (z in that camera's frame)."""
    n, c = _scene_plane()
    k = proj_full[0, view, 1, :3, :3].double()
    e = proj_full[0, view, 0].double()
    rot, t = e[:3, :3], e[:3, 3]
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float64), torch.arange(width, dtype=torch.float64), indexing="ij")
    dirs = rot.T @ torch.linalg.solve(k, torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(3, -1))
    centre = -(rot.T @ t)
    lam = (c - float(n @ centre)) / (n[:, None] * dirs).sum(0)
    pts = centre[:, None] + lam * dirs
    return (rot @ pts + t[:, None])[2].reshape(height, width).float()


def make_scene_images(height: int, width: int, num_views: int, proj_full: torch.Tensor, seed: int = 0,
                      noise: float = 0.01) -> torch.Tensor:
    """[1, N, 3, H, W] renderings of ONE textured plane seen by the rig's cameras (``proj_full`` = the full-resolution
    ``proj_matrices["stage3"]``, [1,N,2,4,4]), plus ``noise`` per-view pixel noise.

    Unlike ``make_images`` the views are photo-consistent with the cameras, as photographs of a scene are: the cost
    volume then has a ridge at the true depth, the regularisation nets regress a piecewise-smooth depth map even with
    random weights, and the per-pixel hypotheses of stages 2/3 (hence the gather pattern of the warp) look like those a
    trained network produces on DTU instead of white noise over the whole depth range."""
    g = torch.Generator().manual_seed(seed)
    n, c = _scene_plane()
    # texture parametrised by the world (x, y) of the surface point, 4 texels per millimetre-ish at DTU scale
    fx = float(proj_full[0, 0, 1, 0, 0])
    mm_per_px = 680.0 / fx
    ext_mm = (width * mm_per_px + 700.0, height * mm_per_px + 700.0)
    tw, th = int(ext_mm[0] / mm_per_px * 0.75), int(ext_mm[1] / mm_per_px * 0.75)
    tex = _fractal_texture(th, tw, g)
    tex = (tex / 4.0 + 0.5).clamp_(0.0, 1.0)
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float64), torch.arange(width, dtype=torch.float64), indexing="ij")
    pix = torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(3, -1)
    views = []
    for v in range(num_views):
        k = proj_full[0, v, 1, :3, :3].double()
        e = proj_full[0, v, 0].double()
        rot, t = e[:3, :3], e[:3, 3]
        dirs = rot.T @ torch.linalg.solve(k, pix)          # ray directions in the world frame
        centre = -(rot.T @ t)
        lam = (c - float(n @ centre)) / (n[:, None] * dirs).sum(0)
        pts = centre[:, None] + lam * dirs                 # [3, H*W] surface points
        gx = (pts[0] / (ext_mm[0] / 2.0)).reshape(1, height, width)
        gy = (pts[1] / (ext_mm[1] / 2.0)).reshape(1, height, width)
        grid = torch.stack([gx, gy], -1).float()
        img = torch.nn.functional.grid_sample(tex[None], grid, mode="bilinear", padding_mode="border", align_corners=False)[0]
        views.append(img + noise * torch.randn(3, height, width, generator=g))
    return torch.stack(views, 0)[None].clamp_(0.0, 1.0).contiguous()


def make_scene_features(height: int, width: int, num_views: int, proj: Dict[str, torch.Tensor], seed: int = 0, noise: float = 0.1,
                        num_stages: int = 3) -> List[Dict[str, torch.Tensor]]:
    """Per-view feature dicts (like FeatureNet's output) that are PHOTO-CONSISTENT with the cameras: every ``stageK`` /
    ``stageK_c`` map is a zero-mean, unit-variance C-channel texture painted on the plane of ``scene_depth`` and seen through
    view v's stage-K camera, plus ``noise`` per-view noise - what a trained FeatureNet computes from photographs of a scene
    (descriptors that agree where the views see the same surface point).  The group-wise correlation then has its ridge at the
    true depth for every pixel and source, which randomly initialised FeatureNet features of rendered images do not give."""
    g = torch.Generator().manual_seed(seed)
    n, c0 = _scene_plane()
    feats: List[Dict[str, torch.Tensor]] = [dict() for _ in range(num_views)]
    for s in range(num_stages):
        scale = 2 ** (num_stages - s - 1)
        h, w = height // scale, width // scale
        c = FEATURE_CHANNELS[s]
        pm = proj["stage%d" % (s + 1)]
        fx = float(pm[0, 0, 1, 0, 0])
        mm_per_px = 680.0 / fx
        ext_mm = (w * mm_per_px + 700.0, h * mm_per_px + 700.0)
        tw, th = int(ext_mm[0] / mm_per_px), int(ext_mm[1] / mm_per_px)
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing="ij")
        pix = torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(3, -1)
        for suffix in ("", "_c"):
            tex = torch.nn.functional.avg_pool2d(torch.randn(1, c, th + 2, tw + 2, generator=g), 3, 1, 0)
            tex = tex / tex.std()
            for v in range(num_views):
                k = pm[0, v, 1, :3, :3].double()
                e = pm[0, v, 0].double()
                rot, t = e[:3, :3], e[:3, 3]
                dirs = rot.T @ torch.linalg.solve(k, pix)
                centre = -(rot.T @ t)
                lam = (c0 - float(n @ centre)) / (n[:, None] * dirs).sum(0)
                pts = centre[:, None] + lam * dirs
                grid = torch.stack([(pts[0] / (ext_mm[0] / 2.0)).reshape(1, h, w), (pts[1] / (ext_mm[1] / 2.0)).reshape(1, h, w)], -1).float()
                f = torch.nn.functional.grid_sample(tex, grid, mode="bilinear", padding_mode="border", align_corners=False)
                feats[v]["stage%d%s" % (s + 1, suffix)] = (f + noise * torch.randn(1, c, h, w, generator=g)).contiguous()
    return feats


def make_stage_features(height: int, width: int, num_views: int, batch: int = 1, seed: int = 0,
                        num_stages: int = 3, structured: bool = True) -> List[Dict[str, torch.Tensor]]:
    """Per-view dicts ``{"stageK": [B,C,h,w], "stageK_c": [B,C,h,w]}`` like FeatureNet's output.

    ``structured=True`` follows the Appendix-D recipe: one smoothed texture,
    cropped with the disparity that a fronto-parallel plane at z0 = 650 induces
    for an x-translated camera, plus noise - this gives a peaked cost volume.
    ``structured=False`` is plain N(0, 1).
    """
    g = torch.Generator().manual_seed(seed)
    feats: List[Dict[str, torch.Tensor]] = [dict() for _ in range(num_views)]
    for s in range(num_stages):
        scale = 2 ** (3 - s - 1)  # stage1 is always 1/4 resolution (reference mvsnet.py:214)
        h, w = height // scale, width // scale
        c = FEATURE_CHANNELS[s]
        for suffix in ("", "_c"):
            key = "stage%d%s" % (s + 1, suffix)
            if not structured:
                for v in range(num_views):
                    feats[v][key] = torch.randn(batch, c, h, w, generator=g)
                continue
            pad = max(8, w // 6)
            base = torch.randn(batch, c, h + 2, w + 2 * pad + 2, generator=g)
            base = 3.0 * torch.nn.functional.avg_pool2d(base, 3, 1, 0)  # [B,c,h,w+2*pad]
            f = 2892.33 * (width / 1600.0) / scale
            for v in range(num_views):
                tx = 0.0 if v == 0 else _RIG_CENTRES[v - 1][0]
                # disparity of a plane at z0 = 650 for cameras converging at z = 680
                off = int(round(f * tx * (1.0 / 650.0 - 1.0 / _LOOK_AT[2])))
                off = max(-pad, min(pad, off))
                crop = base[:, :, :, pad + off: pad + off + w]
                noise = 0.0 if v == 0 else 0.1 * torch.randn(batch, c, h, w, generator=g)
                feats[v][key] = (crop + noise).contiguous()
    return feats


def randomise_regnet_state(state: Dict[str, torch.Tensor], seed: int = 0, prob_gain: float = 4.0,
                           branch_jitter: float = 0.03) -> Dict[str, torch.Tensor]:
    """Non-degenerate parameters for every key of an ``MVSNet.state_dict()``.

    conv weights ~ N(0, 2/fan_in) (stride-2 transposed convs: fan_in / 8 in 3-D, / 4 in 2-D),
    BN running_mean ~ 0.2 N(0,1), running_var ~ U(0.5,1.5), gamma ~ U(0.8,1.2), beta ~ 0.1 N(0,1),
    the ``prob`` convs multiplied by ``prob_gain`` so the softmax over depth is peaked.

    A trained DMVSNet regresses four *nearby* depths per pixel (the dual-depth pairs bracket the surface);
    four independent random nets would disagree by hundreds of millimetres and the (3a-2b ...) extrapolation
    of networks/mvsnet.py:42-45 would throw the refine hypotheses behind the camera, which makes every
    later seam ill-conditioned.  So, like a trained net, the four heads are made similar: ``cosR_huge`` is
    ``cosR_small`` with a relative jitter, and output channel 1 of each ``prob`` is channel 0 with a jitter.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for key in sorted(state.keys()):
        t = state[key]
        if key.endswith("num_batches_tracked"):
            out[key] = torch.zeros_like(t)
        elif key.endswith("running_mean"):
            out[key] = 0.2 * torch.randn(t.shape, generator=g)
        elif key.endswith("running_var"):
            out[key] = 0.5 + torch.rand(t.shape, generator=g)
        elif key.endswith("bn.weight"):
            out[key] = 0.8 + 0.4 * torch.rand(t.shape, generator=g)
        elif key.endswith("bn.bias") or key.endswith(".bias"):
            out[key] = 0.1 * torch.randn(t.shape, generator=g)
        elif key.endswith("weight") and t.dim() >= 4:
            ksz = int(torch.tensor(t.shape[2:]).prod())
            transposed = (".conv7." in key or ".conv9." in key or ".conv11." in key) and "cost_regularization" in key
            if transposed:  # ConvTranspose weight is [Cin, Cout, k...]; each output sees k^n / 2^n taps
                fan_in = t.shape[0] * ksz / float(2 ** (t.dim() - 2))
            else:
                fan_in = t.shape[1] * ksz
            w = torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in)
            if key.endswith("prob.weight"):
                w = w * prob_gain
                w[1] = w[0] * (1.0 + 2.0 * branch_jitter * torch.randn(w[0].shape, generator=g))
            out[key] = w
        else:
            out[key] = t.clone()
    if branch_jitter is not None:
        for key in sorted(out.keys()):
            if ".cosR_huge." in key and not key.endswith("num_batches_tracked"):
                twin = out[key.replace(".cosR_huge.", ".cosR_small.")]
                out[key] = twin * (1.0 + branch_jitter * torch.randn(twin.shape, generator=g))
    return out


def ridge_regnet_state(state: Dict[str, torch.Tensor], seed: int = 0, gain: float = 24.0, perturb: float = 0.05,
                       branch_jitter: float = 0.03) -> Dict[str, torch.Tensor]:
    """Parameters that make the network behave like a TRAINED DMVSNet on photo-consistent inputs, without a checkpoint.

    A trained cost-regularisation net sharpens the ridge the cost volume has at the true depth; softmax + regression then
    return a piecewise-smooth depth map.  Here every regularisation net (both branches, main and refine) computes

        logits = gain * (cost_g0 + cost_g1) / 2  +  perturb-sized contributions of the whole U-Net

    through its outer skip connection: ``conv0`` copies +/- the group mean of the cost into two of its 8 channels (centre tap,
    identity BatchNorm; ReLU keeps the positive / negative part), ``conv11``'s transposed convolution - fed by the randomly
    weighted coarse levels (App. D recipe) - is scaled down to ``perturb``, and ``prob`` reads (ch0 - ch1) * gain at its centre
    tap plus ``perturb``-sized random taps everywhere else.  All layers keep dense, non-degenerate weights and run exactly as
    with a checkpoint; only the values are chosen so that the regressed depth follows the photo-consistency ridge.
    Used by bench.py's headline workload and the free-running full-size parity test: the random-weight recipe regresses a
    white-noise depth map, which no trained network produces, and turns the cascade into a chaotic map at full size."""
    out = randomise_regnet_state(state, seed=seed, branch_jitter=branch_jitter)
    g = torch.Generator().manual_seed(seed + 1000)
    for key in sorted(out.keys()):
        if "cost_regularization" not in key:
            continue
        t = out[key]
        if key.endswith("conv0.conv.weight"):
            w = perturb * t
            mid = tuple(k // 2 for k in t.shape[2:])
            w[(0, slice(None)) + mid] = 0.5
            w[(1, slice(None)) + mid] = -0.5
            out[key] = w
        elif ".conv0.bn." in key:
            out[key] = {"weight": torch.ones_like(t), "bias": torch.zeros_like(t), "running_mean": torch.zeros_like(t),
                        "running_var": torch.ones_like(t) - 1e-5}.get(key.rsplit(".", 1)[1], t)
        elif key.endswith("conv11.bn.weight"):
            out[key] = perturb * t
        elif key.endswith("conv11.bn.bias"):
            out[key] = torch.zeros_like(t)
        elif key.endswith("prob.weight"):
            w = (perturb / 4.0) * t  # randomise_regnet_state scaled these by its prob_gain = 4
            mid = tuple(k // 2 for k in t.shape[2:])
            for o in range(t.shape[0]):
                jit = 1.0 + branch_jitter * float(torch.randn((), generator=g))
                w[(o, 0) + mid] = gain * jit
                w[(o, 1) + mid] = -gain * jit
            out[key] = w
    return out


def calibrate_feature_heads(state: Dict[str, torch.Tensor], imgs: torch.Tensor, crop: Tuple[int, int] = (256, 320)) -> Dict[str, torch.Tensor]:
    """Make a randomly initialised FeatureNet DISCRIMINATIVE on ``imgs`` ([1,N,3,H,W] in [0,1]) the way a trained one is, without a
    checkpoint: the three bias-free output heads (``feature.out1/2/3``, module.py:274-340) are made orthogonal to the mean
    activation vector of their input and scaled to unit output variance, measured on a centre crop of view 0.

    Random heads on post-ReLU activations return feature maps with a large common-mode component; the group-wise correlation
    of two such maps is dominated by it, the cost volume has no ridge, and the ridge-following regularisation nets regress
    noise (24 % mean deviation from the rendered scene; 2 % with calibrated heads at 256 x 320, CPU restatement).  A trained
    FeatureNet's descriptors are centred by training; here the same is obtained from the statistics of the synthetic images.
    Every layer keeps dense random weights and runs exactly as with a checkpoint; workload generator code (plain torch on
    the host), not part of the CUDA path."""
    import torch.nn.functional as F
    st = dict(state)
    h, w = imgs.shape[-2:]
    ch, cw = min(crop[0], h) // 8 * 8, min(crop[1], w) // 8 * 8
    y0, x0 = (h - ch) // 2, (w - cw) // 2
    x = imgs[0, :1, :, y0:y0 + ch, x0:x0 + cw].to("cpu", torch.float32)
    if imgs.dtype == torch.uint8:
        x = x / 255.0
    p = {k[len("feature."):]: v.detach().to("cpu", torch.float32) for k, v in st.items() if k.startswith("feature.")}

    def seq(t, base, specs):
        for i, (stride, pad) in enumerate(specs):
            n = "%s.%d" % (base, i)
            t = F.conv2d(t, p[n + ".conv.weight"], None, stride=stride, padding=pad)
            t = F.relu(F.batch_norm(t, p[n + ".bn.running_mean"], p[n + ".bn.running_var"], p[n + ".bn.weight"], p[n + ".bn.bias"],
                                    False, 0.1, 1e-5))
        return t
    with torch.no_grad():
        c0 = seq(x, "conv0", [(1, 1), (1, 1)])
        c1 = seq(c0, "conv1", [(2, 2), (1, 1), (1, 1)])
        c2 = seq(c1, "conv2", [(2, 2), (1, 1), (1, 1)])
        top2 = F.interpolate(c2, scale_factor=2, mode="nearest") + F.conv2d(c1, p["inner1.weight"], p["inner1.bias"])
        top3 = F.interpolate(top2, scale_factor=2, mode="nearest") + F.conv2d(c0, p["inner2.weight"], p["inner2.bias"])
        for head, inp in (("out1", c2), ("out2", top2), ("out3", top3)):
            wgt = p[head + ".weight"].clone()
            m = inp.mean(dim=(0, 2, 3))
            taps = wgt.shape[2] * wgt.shape[3]
            dc = (wgt.sum(dim=(2, 3)) * m[None]).sum(1)  # response of each filter to the mean input
            wgt = wgt - (dc / (taps * (m * m).sum()))[:, None, None, None] * m[None, :, None, None]
            out = F.conv2d(inp, wgt, None, padding=wgt.shape[-1] // 2)
            wgt = wgt / out.std(dim=(0, 2, 3)).clamp_min(1e-6)[:, None, None, None]
            key = "feature.%s.weight" % head
            st[key] = wgt.to(state[key].device, state[key].dtype)
    return st


def stage_shapes(height: int, width: int, num_stages: int = 3) -> List[Tuple[int, int]]:
    return [(height // 2 ** (3 - s - 1), width // 2 ** (3 - s - 1)) for s in range(num_stages)]
