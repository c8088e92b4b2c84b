"""CPU oracle for the DMVSNet cost-volume hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement, on PyTorch-CPU fp32, of the reference's
algorithm for the path ``BASELINE.json:north_star`` names.  It is the checker
for the CUDA path; nothing under ``dmvsnet_b200/`` may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` leg,
``--impl reference``, and the ``gpu_library_baseline`` leg, which times this
same restatement with its tensors moved to the B200 - the reference's own
PyTorch-CUDA / cuDNN path of SURVEY 8d - as a baseline beside the product,
never inside it) use it.

Parity pinning: the reference ships no tests, fixtures or golden vectors
 - "parity unpinned" by the reference's own tests.  The oracle is
instead pinned against the *live reference itself*: ``tools/make_golden.py``
imports ``/root/reference/networks`` in the build container, checks every
function here against it (bit-exact for the heads and the sampler, <= a few
ulp for warp+corr and the U-Nets) and commits the reference's outputs as
fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the
oracle against those fixtures wherever the suite runs.

Each function cites the reference lines it follows (paths relative to the
reference checkout).  The arithmetic lives in ATen (``grid_sample``, ``conv3d``
...), exactly as in the reference, so the same library calls are used here -
the restatement is of the *algorithm around them*: operation order, layouts,
masks, constants.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm{2,3}d default, networks/module.py:50,144


# --------------------------------------------------------------------------------------
# a3  homography warp                                         networks/module.py:212-251
# --------------------------------------------------------------------------------------
def compose_projection(view_proj: torch.Tensor) -> torch.Tensor:
    """[B,2,4,4] (extrinsic, intrinsic) -> [B,4,4] with the top 3x4 replaced by K @ E[:3,:4].

    networks/mvsnet.py:133-136: the last row of the extrinsic is kept as it is.
    """
    p = view_proj[:, 0].clone()
    p[:, :3, :4] = torch.matmul(view_proj[:, 1, :3, :3], view_proj[:, 0, :3, :4])
    return p


def relative_projection(src_proj: torch.Tensor, ref_proj: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """rot [B,3,3], trans [B,3] of ``P_src @ inv(P_ref)``  (networks/module.py:223-225)."""
    m = torch.matmul(src_proj, torch.inverse(ref_proj))
    return m[:, :3, :3].contiguous(), m[:, :3, 3].contiguous()


def sampling_grid(rot: torch.Tensor, trans: torch.Tensor, hyp: torch.Tensor) -> torch.Tensor:
    """Normalised sampling grid [B,D,H,W,2] for per-pixel hypotheses hyp [B,D,H,W].

    networks/module.py:227-243.  Order of operations is kept: rotate the homogeneous
    pixel, scale by depth, translate, patch exact zeros in Z, divide, normalise.
    """
    b, d, h, w = hyp.shape
    dev = hyp.device  # CPU in every parity test; bench.py's gpu_library_baseline leg runs the same code on the B200 through cuDNN
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dev), torch.arange(w, dtype=torch.float32, device=dev), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=dev)))  # [3,HW]
    rotated = torch.matmul(rot, pix.unsqueeze(0).expand(b, -1, -1))  # [B,3,HW]
    pts = rotated.unsqueeze(2) * hyp.reshape(b, 1, d, h * w) + trans.reshape(b, 3, 1, 1)
    z = pts[:, 2]
    z = torch.where(z == 0, z + 1e-5, z)
    u = pts[:, 0] / z
    v = pts[:, 1] / z
    un = u / ((w - 1) / 2) - 1
    vn = v / ((h - 1) / 2) - 1
    return torch.stack((un, vn), dim=-1).reshape(b, d, h, w, 2)


def warp_features(src: torch.Tensor, rot: torch.Tensor, trans: torch.Tensor, hyp: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> [B,C,D,H,W]; bilinear, zeros padding, align_corners=True (module.py:247-249)."""
    b, c, h, w = src.shape
    d = hyp.shape[1]
    grid = sampling_grid(rot, trans, hyp)
    out = F.grid_sample(src, grid.reshape(b, d * h, w, 2), mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.reshape(b, c, d, h, w)


# --------------------------------------------------------------------------------------
# a2  group-wise correlation summed over source views          networks/mvsnet.py:111-153
# --------------------------------------------------------------------------------------
def warp_corr(features: Sequence[torch.Tensor], proj_matrices: torch.Tensor, hyp: torch.Tensor) -> torch.Tensor:
    """features: N x [B,C,h,w] (ref first); proj_matrices [B,N,2,4,4]; hyp [B,D,h,w] -> [B,2,D,h,w].

    Two interleaved channel groups (group g = channels 2j+g), mean over C/2 channels
    (mvsnet.py:139), plain sum over source views, no division (mvsnet.py:141-146).
    """
    ref = features[0]
    b, c, h, w = ref.shape
    ref_proj = compose_projection(proj_matrices[:, 0])
    ref_g = ref.reshape(b, c // 2, 2, 1, h, w)
    total = None
    for v in range(1, len(features)):
        rot, trans = relative_projection(compose_projection(proj_matrices[:, v]), ref_proj)
        warped = warp_features(features[v], rot, trans, hyp)
        sim = (warped.reshape(b, c // 2, 2, -1, h, w) * ref_g).mean(1)
        total = sim if total is None else total + sim
    return total


# --------------------------------------------------------------------------------------
# a9 / a4 / a5  conv blocks and the two U-Nets                 networks/module.py:28-208,342-436
# --------------------------------------------------------------------------------------
def _bn(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"],
                        p[prefix + ".weight"], p[prefix + ".bias"], False, 0.1, BN_EPS)


def _block(x: torch.Tensor, p: Dict[str, torch.Tensor], name: str, *, dims: int, stride: int = 1,
           transposed: bool = False, padding: int = 1) -> torch.Tensor:
    """conv (no bias) -> eval BatchNorm -> ReLU   (module.py:57-63,102-111,151-157,196-202)."""
    wgt = p[name + ".conv.weight"]
    if transposed:
        fn = F.conv_transpose3d if dims == 3 else F.conv_transpose2d
        y = fn(x, wgt, None, stride=stride, padding=padding, output_padding=1)
        if dims == 2:  # Deconv2d crops to exactly twice the input, module.py:104-106
            y = y[:, :, : 2 * x.shape[2], : 2 * x.shape[3]].contiguous()
    else:
        fn = F.conv3d if dims == 3 else F.conv2d
        y = fn(x, wgt, None, stride=stride, padding=padding)
    return F.relu(_bn(y, p, name + ".bn"))


def _sub(p: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


def regnet_branch(x: torch.Tensor, p: Dict[str, torch.Tensor]) -> torch.Tensor:
    """One ``CostRegNet_part`` (module.py:358-398).  x [B,2,D,h,w] -> [B,2,D,h,w]."""
    c0 = _block(x, p, "conv0", dims=3)
    c2 = _block(_block(c0, p, "conv1", dims=3, stride=2), p, "conv2", dims=3)
    c4 = _block(_block(c2, p, "conv3", dims=3, stride=2), p, "conv4", dims=3)
    y = _block(_block(c4, p, "conv5", dims=3, stride=2), p, "conv6", dims=3)
    y = c4 + _block(y, p, "conv7", dims=3, stride=2, transposed=True)
    y = c2 + _block(y, p, "conv9", dims=3, stride=2, transposed=True)
    y = c0 + _block(y, p, "conv11", dims=3, stride=2, transposed=True)
    return F.conv3d(y, p["prob.weight"], None, stride=1, padding=1)


def regnet_branch_refine(x: torch.Tensor, p: Dict[str, torch.Tensor]) -> torch.Tensor:
    """One ``CostRegNet_part_refine`` (module.py:400-436): D = 4 -> 2 -> 1, 2-D bottleneck."""
    c0 = _block(x, p, "conv0", dims=3)
    c2 = _block(_block(c0, p, "conv1", dims=3, stride=2), p, "conv2", dims=3)
    c4 = _block(_block(c2, p, "conv3", dims=3, stride=2), p, "conv4", dims=3).squeeze(2)
    y = _block(_block(c4, p, "conv5", dims=2, stride=2), p, "conv6", dims=2)
    y = c4 + _block(y, p, "conv7", dims=2, stride=2, transposed=True)
    y = y.unsqueeze(2)
    y = c2 + _block(y, p, "conv9", dims=3, stride=2, transposed=True)
    y = c0 + _block(y, p, "conv11", dims=3, stride=2, transposed=True)
    return F.conv3d(y, p["prob.weight"], None, stride=1, padding=1)


def regnet(x: torch.Tensor, p: Dict[str, torch.Tensor], refine: bool = False) -> torch.Tensor:
    """``CostRegNet`` / ``CostRegNet_refine`` (module.py:342-357): two branches, concatenated -> 4 channels."""
    fn = regnet_branch_refine if refine else regnet_branch
    return torch.cat((fn(x, _sub(p, "cosR_small.")), fn(x, _sub(p, "cosR_huge."))), dim=1)


# --------------------------------------------------------------------------------------
# a6  dual-depth head                                          networks/mvsnet.py:15-66
# --------------------------------------------------------------------------------------
def _confidence(d4: torch.Tensor, interval: torch.Tensor) -> torch.Tensor:
    spread = d4.var(1, unbiased=False).sqrt()
    return 2 * (torch.sigmoid(interval / (spread + 1e-5)) - 0.5)


def _stack6(lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    return torch.stack((3 * lo - 2 * hi, 2 * lo - hi, lo, hi, 2 * hi - lo, 3 * hi - 2 * lo), 1)


def depth_head(logits: torch.Tensor, hyp: torch.Tensor, interval: torch.Tensor) -> Dict[str, torch.Tensor]:
    """logits [B,4,D,h,w], hyp [B,D,h,w] -> prob volume, 4 regressed depths, 4 refine hypotheses, confidence.

    Row class r = y % 4 picks (small | huge) x (as is | doubled range); column class x % 2 together
    with the row parity picks the low window stack6[0:4] or the high window stack6[2:6]
    (mvsnet.py:33-56).
    """
    prob = F.softmax(logits, dim=2)
    d4 = torch.sum(prob * hyp.unsqueeze(1), dim=2)  # module.py:454-460
    b, _, h, w = d4.shape
    s_lo, s_hi = d4[:, 0:2].min(1)[0], d4[:, 0:2].max(1)[0]
    g_lo, g_hi = d4[:, 2:4].min(1)[0], d4[:, 2:4].max(1)[0]
    g_lo2, g_hi2 = 2 * g_lo - g_hi, 2 * g_hi - g_lo
    s_lo2, s_hi2 = 2 * s_lo - s_hi, 2 * s_hi - s_lo
    stacks = [_stack6(s_lo, s_hi), _stack6(g_lo, g_hi), _stack6(s_lo2, s_hi2), _stack6(g_lo2, g_hi2)]
    rows = (torch.arange(h, device=d4.device) % 4).reshape(1, 1, h, 1)
    cols = (torch.arange(w, device=d4.device) % 2).reshape(1, 1, 1, w)
    nxt = torch.zeros_like(d4)
    for r in range(4):
        # rows 0,2 (small): even column -> low window; rows 1,3 (huge): even column -> high window
        low_on_even_col = (r % 2 == 0)
        for c in range(2):
            low = (c == 0) == low_on_even_col
            window = stacks[r][:, 0:4] if low else stacks[r][:, 2:6]
            nxt = torch.where((rows == r) & (cols == c), window, nxt)
    return {"photometric_confidence": _confidence(d4, interval), "prob_volume": prob, "depth_sub_plus": d4,
            "depth_values_c": nxt, "depth_values": hyp, "interval": interval}


# --------------------------------------------------------------------------------------
# a7  refine head                                               networks/mvsnet.py:67-100
# --------------------------------------------------------------------------------------
def refine_head(logits: torch.Tensor, hyp_c: torch.Tensor, interval: torch.Tensor, alpha: float = 5) -> Dict[str, torch.Tensor]:
    prob = F.softmax(logits * alpha, dim=2)
    d4 = torch.sum(prob * hyp_c.unsqueeze(1), dim=2)
    b, _, h, w = d4.shape
    s_lo, s_hi = d4[:, 0:2].min(1)[0], d4[:, 0:2].max(1)[0]
    g_lo, g_hi = d4[:, 2:4].min(1)[0], d4[:, 2:4].max(1)[0]
    ry = (torch.arange(h, device=d4.device) % 2).reshape(1, h, 1)
    cx = (torch.arange(w, device=d4.device) % 2).reshape(1, 1, w)
    depth = torch.where(ry == 0, torch.where(cx == 0, s_lo, s_hi), torch.where(cx == 0, g_hi, g_lo))
    return {"depth": depth, "photometric_confidence_refine": _confidence(d4, interval), "depth_sub_plus_refine": d4}


# --------------------------------------------------------------------------------------
# a1  hypothesis sampler + upsample                networks/module.py:476-649, mvsnet.py:224-233
# --------------------------------------------------------------------------------------
def depth_hypotheses(last_depth: torch.Tensor, ndepth: int, interval_pixel, shape: Optional[Sequence[int]] = None,
                     inverse: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (samples [B,D,h,w] at the resolution of ``last_depth`` (or ``shape`` at stage 0), interval).

    Checkerboard parity ``even = (y + x) % 2 == 0`` selects the "n" (shifted down) or "p"
    (shifted up) range.  Stage 0 reads only the first and last entry of ``last_depth`` [B,Nd]
    and ignores ``interval_pixel`` (module.py:560-579, 598-634).
    """
    if last_depth.dim() == 2:
        h, w = int(shape[0]), int(shape[1])
        lo, hi = last_depth[:, 0], last_depth[:, -1]
        step = (hi - lo) / (ndepth - 1)
        si = step[0]  # batch 0 only, module.py:564,603
        k = torch.arange(ndepth, dtype=last_depth.dtype, device=last_depth.device).reshape(1, -1)
        if not inverse:
            planes = lo.unsqueeze(1) + k * step.unsqueeze(1)
            planes_n, planes_p = planes - si, planes + si
        else:
            def inv_planes(a, b_):
                return 1 / torch.stack([torch.linspace(float(1 / x), float(1 / y), ndepth, device=last_depth.device) for x, y in zip(a, b_)])
            # the interval is recomputed after the shift and comes out unchanged (module.py:606-621)
            planes_n = inv_planes(lo - si, hi - si)
            si2 = (((hi - si) - (lo - si)) / (ndepth - 1))[0]
            planes_p = inv_planes(lo + si2, hi + si2)
            si = ((((hi + si2) - (lo + si2)) / (ndepth - 1))[0])
        ys = torch.arange(h, device=last_depth.device).reshape(1, 1, h, 1)
        xs = torch.arange(w, device=last_depth.device).reshape(1, 1, 1, w)
        even = ((ys + xs) % 2) == 0
        samples = torch.where(even, planes_n.reshape(-1, ndepth, 1, 1), planes_p.reshape(-1, ndepth, 1, 1))
        return samples.float().contiguous(), si.float() if inverse else si
    b, h, w = last_depth.shape
    k = torch.arange(ndepth, dtype=last_depth.dtype, device=last_depth.device).reshape(1, -1, 1, 1)

    def ranged(lo_off, hi_off):
        lo = last_depth - lo_off / 2 * interval_pixel
        hi = last_depth + hi_off / 2 * interval_pixel
        if inverse:
            ilo, ihi = 1 / lo, 1 / hi
            return 1 / (ilo.unsqueeze(1) + k * ((ihi - ilo) / (ndepth - 1)).unsqueeze(1))
        return lo.unsqueeze(1) + k * ((hi - lo) / (ndepth - 1)).unsqueeze(1)

    samples_n = ranged(ndepth + 2, ndepth - 2)   # module.py:476-491 / 540-554
    samples_p = ranged(ndepth - 2, ndepth + 2)   # module.py:492-507 / 525-539
    ys = torch.arange(h, device=last_depth.device).reshape(1, 1, h, 1)
    xs = torch.arange(w, device=last_depth.device).reshape(1, 1, 1, w)
    even = ((ys + xs) % 2) == 0
    interval = (ndepth * interval_pixel) / (ndepth - 1)
    samples = torch.where(even, samples_n, samples_p)
    if inverse:
        return samples.float(), interval.float() if torch.is_tensor(interval) else interval
    return samples, interval


def upsample_hypotheses(samples: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    """mvsnet.py:232-233: bilinear, align_corners=False."""
    return F.interpolate(samples, [int(shape[0]), int(shape[1])], mode="bilinear", align_corners=False)


# --------------------------------------------------------------------------------------
# N1  FeatureNet (above the hot path; needed for full-forward scope)   module.py:274-340
# --------------------------------------------------------------------------------------
def feature_net(img: torch.Tensor, p: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    def seq(x, base, specs):
        for i, (stride, pad) in enumerate(specs):
            name = "%s.%d" % (base, i)
            y = F.conv2d(x, p[name + ".conv.weight"], None, stride=stride, padding=pad)
            x = F.relu(_bn(y, p, name + ".bn"))
        return x

    c0 = seq(img, "conv0", [(1, 1), (1, 1)])
    c1 = seq(c0, "conv1", [(2, 2), (1, 1), (1, 1)])
    c2 = seq(c1, "conv2", [(2, 2), (1, 1), (1, 1)])
    out = {}

    def split(name, t):
        half = t.shape[1] // 2
        out[name], out[name + "_c"] = t[:, :half], t[:, half:]

    split("stage1", F.conv2d(c2, p["out1.weight"]))
    top = F.interpolate(c2, scale_factor=2, mode="nearest") + F.conv2d(c1, p["inner1.weight"], p["inner1.bias"])
    split("stage2", F.conv2d(top, p["out2.weight"], None, padding=1))
    top = F.interpolate(top, scale_factor=2, mode="nearest") + F.conv2d(c0, p["inner2.weight"], p["inner2.bias"])
    split("stage3", F.conv2d(top, p["out3.weight"], None, padding=1))
    return out


# --------------------------------------------------------------------------------------
# a8  cascade driver                                            networks/mvsnet.py:188-260
# --------------------------------------------------------------------------------------
def cascade_forward(features: List[Dict[str, torch.Tensor]], proj_matrices: Dict[str, torch.Tensor],
                    depth_values: torch.Tensor, state: Dict[str, torch.Tensor], ndepths: Sequence[int],
                    ratios: Sequence[float], inverse_depth: bool, image_hw: Sequence[int],
                    keep_seams: bool = False) -> Dict[str, object]:
    """The stage loop with features precomputed (bench scope H)."""
    depth_interval = (depth_values[0, -1] - depth_values[0, 0]) / depth_values.size(1)
    outputs: Dict[str, object] = {}
    last = depth_values
    for s in range(len(ndepths)):
        name = "stage%d" % (s + 1)
        scale = 2 ** (3 - s - 1)
        shape = [image_hw[0] // scale, image_hw[1] // scale]
        hyp, interval = depth_hypotheses(last, ndepths[s], ratios[s] * depth_interval, shape, inverse_depth)
        if s > 0:
            hyp = upsample_hypotheses(hyp, shape)
        pm = proj_matrices[name]
        cost = warp_corr([f[name] for f in features], pm, hyp)
        logits = regnet(cost, _sub(state, "cost_regularization.%d." % s))
        main = depth_head(logits, hyp, interval)
        hyp_c = main["depth_values_c"]
        cost_c = warp_corr([f[name + "_c"] for f in features], pm, hyp_c)
        logits_c = regnet(cost_c, _sub(state, "cost_regularization_refine.%d." % s), refine=True)
        ref = refine_head(logits_c, hyp_c, interval)
        merged = {**ref, **main}
        if keep_seams:
            merged.update({"_cost": cost, "_logits": logits, "_cost_c": cost_c, "_logits_c": logits_c})
        last = merged["depth"].detach()
        outputs[name] = merged
        outputs.update(merged)
    return outputs


def mvsnet_forward(imgs: torch.Tensor, proj_matrices: Dict[str, torch.Tensor], depth_values: torch.Tensor,
                   state: Dict[str, torch.Tensor], ndepths: Sequence[int], ratios: Sequence[float],
                   inverse_depth: bool = False, keep_seams: bool = False) -> Dict[str, object]:
    """Full ``MVSNet.forward`` (bench scope F): FeatureNet per view, then the cascade."""
    fp = _sub(state, "feature.")
    features = [feature_net(imgs[:, v], fp) for v in range(imgs.shape[1])]
    return cascade_forward(features, proj_matrices, depth_values, state, ndepths, ratios, inverse_depth,
                           imgs.shape[-2:], keep_seams)
