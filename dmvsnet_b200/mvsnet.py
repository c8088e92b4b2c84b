"""Host-side mirror of the reference's ``networks/mvsnet.py``: the drop-in boundary.

``MVSNet(ndepths, depth_interval_ratio, ...)`` keeps the reference constructor signature
(networks/mvsnet.py:157), ``forward(imgs, proj_matrices, depth_values) -> dict`` (mvsnet.py:188), the 787
``state_dict`` keys and every output key (``depth``, ``photometric_confidence``, ``prob_volume``,
``depth_sub_plus``, ``depth_values_c``, ``stage1..3`` ...), so ``model.py`` / ``loss.py`` of the reference
work against it unchanged.  The stage loop runs on hand-written sm_100a kernels:

    per stage:  S1 hypotheses -> W1 warp+corr -> R1 U-Nets -> E1 head -> W1 (D=4, *_c features) -> R1 refine -> E2

Inference only (eval-mode BatchNorm); training raises.  CUDA only; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .module import CostRegNet, CostRegNet_refine, FeatureNet, _require_inference

__all__ = ["MVSNet", "CostAgg", "DepthNet", "Align_Corners_Range"]

Align_Corners_Range = False  # mvsnet.py:8; the fused sampler implements align_corners=False


def _as_float_images(imgs: torch.Tensor) -> torch.Tensor:
    """uint8 photographs -> the fp32 [0,1] images the network is defined on: x / 255 in IEEE fp32, as the reference's loaders."""
    if imgs.dtype == torch.uint8:
        # a TENSOR divisor: torch turns division by a Python scalar into a multiplication by 1/255 on CUDA, which is not the
        # correctly rounded quotient numpy computes on the host
        return imgs.to(torch.float32) / torch.full((), 255.0, device=imgs.device)
    return imgs


def _check_finite(flag: torch.Tensor) -> None:
    if not bool(flag[0]):
        raise FloatingPointError(
            "dmvsnet_b200: the regressed depth map is not finite.  Either the inputs are (NaN / Inf depth range or images) or an "
            "activation left the fp16 hi/lo range of the tensor engine (|x| > 65504, ops.FP16_SPLIT_MAX): run that network with "
            "ops.DEFAULT_ENGINE = 'fp32'.")


class DepthNet(nn.Module):
    """Dual-depth heads.  reference networks/mvsnet.py:11-100."""

    def __init__(self, mode="regression"):
        super().__init__()
        self.mode = mode
        self.return_prob_volume = True  # the reference always returns it; turn off to save 4*D*h*w*4 bytes of writes

    def forward(self, cost_reg, depth_values, num_depth, interval, prob_volume_init=None, stage=0):
        prob, d4, hyp_c, conf = ops.depth_head(cost_reg, depth_values, interval, want_prob=self.return_prob_volume)
        out = {"photometric_confidence": conf, "depth_sub_plus": d4, "depth_values_c": hyp_c, "depth_values": depth_values,
               "interval": interval}
        if prob is not None:
            out["prob_volume"] = prob
        return out

    def refine(self, cost_reg, depth_values, num_depth, interval, alpha=5):
        depth, conf, d4 = ops.refine_head(cost_reg, depth_values, interval, alpha)
        return {"depth": depth, "photometric_confidence_refine": conf, "depth_sub_plus_refine": d4}


class CostAgg(nn.Module):
    """Group-wise correlation cost volume over the source views.  reference networks/mvsnet.py:102-153."""

    def __init__(self, mode="variance", in_channels=None):
        super().__init__()
        assert mode in ("variance", "adaptive"), "Don't support {}!".format(mode)
        if mode == "adaptive":
            raise NotImplementedError("agg_mode='adaptive' only adds an unused weight_net in the reference (mvsnet.py:107-108)")
        self.mode = mode

    def forward(self, features, proj_matrices, depth_values, stage_idx, rt: Optional[torch.Tensor] = None):
        """features: [ref, src1, ...] each [B,C,h,w]; proj_matrices [B,N,2,4,4]; depth_values [B,D,h,w] -> [B,2,D,h,w].

        ``rt`` lets the cascade pass homographies it has already computed for this stage."""
        if rt is None:
            rt = ops.relative_projections(proj_matrices).to(features[0].device, non_blocking=True)
        if torch.is_grad_enabled() and any(f.requires_grad for f in features):
            # differentiable like the reference's (gradients to the feature maps; the grid is built under no_grad, module.py:222)
            return ops.warp_corr_autograd(features, rt, depth_values)
        return ops.warp_corr(features, rt, depth_values)

    def forward_fused(self, features, depth_values, rt, want_f32=False, layout=None, coherent=False):
        """Cascade-internal: W1 writes the cost volume directly in the cell layout the tensor engine's first conv reads by
        TMA (no fp32 round trip through HBM unless ``want_f32``).  Returns (cost_or_None, cells)."""
        return ops.warp_corr(features, rt, depth_values, want_f32=want_f32, want_cells=True, layout=layout, coherent=coherent)


class MVSNet(nn.Module):
    def __init__(self, ndepths, depth_interval_ratio, cr_base_chs=None, fea_mode="fpn", agg_mode="variance",
                 depth_mode="regression", winner_take_all_to_generate_depth=True, inverse_depth=False):
        super().__init__()
        if cr_base_chs is None:
            cr_base_chs = [8] * len(ndepths)
        assert len(ndepths) == len(depth_interval_ratio)
        self.ndepths = ndepths
        self.depth_interval_ratio = depth_interval_ratio
        self.fea_mode = fea_mode
        self.cr_base_chs = cr_base_chs
        self.num_stage = len(ndepths)
        self.inverse_depth = inverse_depth
        # infer_many: run FeatureNet of item k+1 on its own stream beside the cascade of item k
        self.overlap_features = True
        # W1 source-map precision.  "fp16": the source views' feature maps are rounded to fp16 once and W1 runs the TMA-staged
        # kernel on them (dmvs_warp_corr_h16_f32: half the bytes through the SMs' shared-memory pipe, 2x faster; regressed depth
        # within 3.3e-4 of the fp32 path on a network that behaves like a trained one, contract 1e-3).  "fp32": the exact kernels
        # (cost volume within 2e-6 of the reference's).
        self.w1_precision = "fp16"

        self.feature = FeatureNet(base_channels=8, stride=4, num_stage=self.num_stage, mode=self.fea_mode)
        self.cost_aggregation = CostAgg(agg_mode, self.feature.out_channels)
        self.cost_regularization = nn.ModuleList(
            [CostRegNet(in_channels=2, base_channels=self.cr_base_chs[i], stage=i) for i in range(self.num_stage)])
        self.cost_regularization_refine = nn.ModuleList(
            [CostRegNet_refine(in_channels=2, base_channels=self.cr_base_chs[i], stage=i) for i in range(self.num_stage)])
        self.DepthNet = DepthNet(depth_mode)

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_copy_stream", "_feat_stream"):  # CUDA streams are per process / device: recreated on demand
            state.pop(k, None)
        return state

    # ------------------------------------------------------------------ the hot path
    def cascade(self, features: Sequence[Dict[str, torch.Tensor]], proj_matrices: Dict[str, torch.Tensor],
                depth_values: torch.Tensor, image_hw: Sequence[int], keep_seams: bool = False,
                rts: Optional[Sequence[torch.Tensor]] = None, branch_group=None) -> Dict[str, object]:
        """``_cascade`` with the feature maps' device made current (the native launches go to the current device)."""
        with torch.cuda.device(features[0]["stage1"].device):
            return self._cascade(features, proj_matrices, depth_values, image_hw, keep_seams, rts, branch_group)

    def _cascade(self, features: Sequence[Dict[str, torch.Tensor]], proj_matrices: Dict[str, torch.Tensor],
                 depth_values: torch.Tensor, image_hw: Sequence[int], keep_seams: bool = False,
                 rts: Optional[Sequence[torch.Tensor]] = None, branch_group=None) -> Dict[str, object]:
        """The stage loop (reference mvsnet.py:208-258) on precomputed per-view feature dicts.

        ``keep_seams`` additionally returns the cost volumes and logits (``_cost``, ``_logits``, ``_cost_c``,
        ``_logits_c``) per stage, for the parity tests.  ``rts`` overrides the per-stage homographies
        ([B,N-1,12] each, see ops.relative_projections): fp32 ``inverse(P_ref)`` is ill-conditioned (entries ~1e5),
        two hosts' LAPACKs differ by ~1e-4 relative in H, so cross-machine fixtures carry the reference's own H.
        ``branch_group``: a 2-rank ``torch.distributed`` group holding the same inputs - single-view latency mode: every
        regularisation net runs one branch per rank (parallel.regnet_branch_sharded), everything else is replicated."""
        _require_inference(self)
        dev = features[0]["stage1"].device
        # K1: all homographies up front, on the host, exactly as the reference computes them; one small upload.
        if rts is None:
            rts = [ops.relative_projections(proj_matrices["stage%d" % (s + 1)]) for s in range(self.num_stage)]
        rts = [r.to(dev, torch.float32, non_blocking=True) for r in rts]
        depth_values = depth_values.to(dev, torch.float32)
        depth_interval = (depth_values[0, -1] - depth_values[0, 0]) / depth_values.size(1)  # batch 0 only, mvsnet.py:196
        outputs: Dict[str, object] = {}
        last_depth = None
        for s in range(self.num_stage):
            name = "stage%d" % (s + 1)
            scale = 2 ** (3 - s - 1)
            shape = [int(image_hw[0]) // scale, int(image_hw[1]) // scale]
            if s == 0:
                hyp, interval = ops.hypotheses_first(depth_values, self.ndepths[s], shape, self.inverse_depth)
            else:
                hyp, interval = ops.hypotheses_next(last_depth.detach(), self.ndepths[s],
                                                    self.depth_interval_ratio[s] * depth_interval, shape, self.inverse_depth)
            fused = ops.DEFAULT_ENGINE == "tensor"
            half = self.w1_precision == "fp16"
            w1_layout = "h16" if half else None

            def maps(key):
                return [features[0][key]] + [(f.get(key + "_h16", f[key]) if half else f[key]) for f in features[1:]]
            # W1 kernel choice (ops.W1_LAYOUT = "auto"): stage-1 planes come from the sampler (two parity classes of
            # fronto-parallel planes), so a pixel tile's source footprint is a small box -> TMA-staged kernel; every later
            # pass has per-pixel hypotheses regressed by the previous pass, whose roughness the channel-last gather does
            # not care about (profiles/, tools/bench_w1.py)
            if fused:
                cost, cells = self.cost_aggregation.forward_fused(maps(name), hyp, rts[s], want_f32=keep_seams, layout=w1_layout,
                                                                  coherent=(s == 0))
            else:
                cost, cells = ops.warp_corr(maps(name), rts[s], hyp, layout=w1_layout, coherent=(s == 0)), None
            logits = self.cost_regularization[s](cost, cost_cells=cells, branch_group=branch_group)
            stage_out = self.DepthNet(logits, hyp, num_depth=self.ndepths[s], interval=interval, stage=s)
            seams = {"_cost": cost, "_logits": logits} if keep_seams else {}
            del cost, logits, cells
            hyp_c = stage_out["depth_values_c"]
            if fused:
                cost_c, cells_c = self.cost_aggregation.forward_fused(maps(name + "_c"), hyp_c, rts[s], want_f32=keep_seams,
                                                                      layout=w1_layout)
            else:
                cost_c, cells_c = ops.warp_corr(maps(name + "_c"), rts[s], hyp_c, layout=w1_layout), None
            logits_c = self.cost_regularization_refine[s](cost_c, cost_cells=cells_c, branch_group=branch_group)
            refine_out = self.DepthNet.refine(logits_c, hyp_c, num_depth=4, interval=interval)
            if keep_seams:
                seams.update({"_cost_c": cost_c, "_logits_c": logits_c})
            stage_out = {**refine_out, **stage_out, **seams}
            last_depth = stage_out["depth"]
            outputs[name] = stage_out
            outputs.update(stage_out)
        return outputs

    def forward(self, imgs, proj_matrices, depth_values):
        """imgs [B,N,3,H,W], proj_matrices {"stageK": [B,N,2,4,4]}, depth_values [B,Nd] -> dict (mvsnet.py:188-260)."""
        _require_inference(self)
        if not imgs.is_cuda:
            raise RuntimeError("dmvsnet_b200.MVSNet runs on CUDA (sm_100a) only; use MVSNet.infer() for host buffers")
        features = self.extract_features(imgs)
        return self.cascade(features, proj_matrices, depth_values, imgs.shape[-2:])

    def extract_features(self, imgs: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
        """FeatureNet on all views in ONE batched call (the reference loops over views, mvsnet.py:199-202; same per-view
        arithmetic, but 5 small cuDNN launches per layer become one: 39 -> 23 ms at DTU size on B200).  Returns the
        reference's per-view list of dicts; every entry is a strided view into the batched output (no copies)."""
        b, n = imgs.shape[0], imgs.shape[1]
        want16 = self.w1_precision == "fp16" and imgs.is_cuda and n > 1
        self.feature.emit_f16 = want16  # the tensor heads' epilogues write the fp16 source maps themselves
        try:
            out = self.feature(imgs.reshape(b * n, *imgs.shape[2:]))
        finally:
            self.feature.emit_f16 = False
        halves = {k[:-4]: out.pop(k) for k in [k for k in out if k.endswith("_h16")]}

        def view_of(t, v):
            return t.view(b, n, *t.shape[1:])[:, v]
        views = [{k: view_of(t, v) for k, t in out.items()} for v in range(n)]
        if want16:
            for k, t in out.items():  # maps without an emitted copy: one conversion launch per map for all views
                half = halves[k] if k in halves else ops.features_nhwc_f16(t)
                for v in range(1, n):
                    views[v][k + "_h16"] = half.view_of(b, n, v)
        return views

    def add_half_features(self, features: List[Dict[str, torch.Tensor]]) -> List[Dict[str, torch.Tensor]]:
        """Attach the fp16 channel-last copy W1's staged kernel gathers from (``<key>_h16``, ops.HalfFeatures) to every
        SOURCE view's maps (view 0 stays fp32: it is the reference view of the cost volume).  In place; returns ``features``."""
        for f in features[1:]:
            for key in [k for k in f if not k.endswith("_h16")]:
                if key + "_h16" not in f:
                    f[key + "_h16"] = ops.features_nhwc_f16(f[key])
        return features

    # ------------------------------------------------------------------ host-buffer entry
    # views per H2D / FeatureNet group in infer(): the copy of group k+1 (side stream) runs under FeatureNet of group k
    infer_view_groups = 3

    @torch.no_grad()
    def infer(self, imgs: torch.Tensor, proj_matrices: Dict[str, torch.Tensor], depth_values: torch.Tensor,
              keys: Sequence[str] = ("depth", "photometric_confidence")) -> Dict[str, torch.Tensor]:
        """Host tensors in, host tensors out: H2D of the images, forward, D2H of ``keys`` only.

        ``imgs`` may be uint8 [B,N,3,H,W] (the decoded photographs, 0..255): they cross PCIe as they are (a quarter of the
        bytes) and are scaled on the device by the same IEEE fp32 division by 255 the reference's loaders do on the host
        (datasets/general_eval.py:161 ``np.array(img, dtype=np.float32) / 255.``) - identical pixel values.

        What ``Model.test`` does around the network (tools.tocuda model.py:333 / tensor2numpy model.py:347), minus the
        D2H of the three probability volumes nobody reads at test time.  The views are uploaded in groups on a side
        stream so that FeatureNet (per-view arithmetic, mvsnet.py:199-202) starts on the first group while the rest is
        still crossing PCIe."""
        _require_inference(self)
        dev = next(self.parameters()).device
        with torch.cuda.device(dev):
            return self._infer(imgs, proj_matrices, depth_values, keys, dev)

    def _infer(self, imgs, proj_matrices, depth_values, keys, dev):
        if imgs.is_cuda:
            out = self.forward(_as_float_images(imgs), proj_matrices, depth_values)
        else:
            src = imgs if imgs.is_pinned() else imgs.pin_memory()
            b, n = src.shape[0], src.shape[1]
            main = torch.cuda.current_stream(dev)
            side = getattr(self, "_copy_stream", None)
            if side is None or side.device != dev:
                side = self._copy_stream = torch.cuda.Stream(dev)
            per = max(1, -(-n // max(1, int(self.infer_view_groups))))
            side.wait_stream(main)
            chunks = []
            for lo in range(0, n, per):
                with torch.cuda.stream(side):
                    part = src[:, lo:lo + per].to(dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(side)
                chunks.append((lo, part, ev))
            feats: List[Dict[str, torch.Tensor]] = []
            for lo, part, ev in chunks:
                main.wait_event(ev)
                part.record_stream(main)
                feats.extend(self.extract_features(_as_float_images(part)))
            out = self.cascade(feats, proj_matrices, depth_values, imgs.shape[-2:])
        host = {}
        finite = torch.empty(1, dtype=torch.bool, pin_memory=True)
        finite.copy_(torch.isfinite(out["depth"]).all().reshape(1), non_blocking=True)
        for k in keys:
            buf = torch.empty(out[k].shape, dtype=out[k].dtype, pin_memory=True)
            buf.copy_(out[k], non_blocking=True)
            host[k] = buf
        torch.cuda.current_stream().synchronize()
        _check_finite(finite)
        return host

    @torch.no_grad()
    def infer_many(self, inputs, keys: Sequence[str] = ("depth", "photometric_confidence")):
        """Generator over ``(imgs, proj_matrices, depth_values)`` host triples -> host result dicts, in order.

        The production loop of ``Model.test`` (model.py:323-380: one reference view after the other) as a three-stage
        pipeline: the images of item k+1 cross PCIe on a side stream while item k computes, and the results of item k
        are copied back while item k+1 computes.  Every item still pays its own H2D and D2H; only their latency is
        hidden.  Results are yielded one item late (after their D2H has completed)."""
        _require_inference(self)
        dev = next(self.parameters()).device
        if dev.index != torch.cuda.current_device():
            # a generator cannot hold a device context across its yields without leaking it to the consumer
            raise RuntimeError("infer_many: make the model's device current first (torch.cuda.set_device(%d))" % dev.index)
        main = torch.cuda.current_stream(dev)
        side = getattr(self, "_copy_stream", None)
        if side is None or side.device != dev:
            side = self._copy_stream = torch.cuda.Stream(dev)

        def upload(item):
            imgs, proj, dv = item
            src = imgs if imgs.is_pinned() else imgs.pin_memory()
            with torch.cuda.stream(side):  # side-stream pool; record_stream(main) below keeps recycling safe
                dimgs = src.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            return dimgs, proj, dv, ev

        feat_stream = None
        if self.overlap_features:
            feat_stream = getattr(self, "_feat_stream", None)
            if feat_stream is None or feat_stream.device != dev:
                feat_stream = self._feat_stream = torch.cuda.Stream(dev)

        def features_of(up):
            """FeatureNet of an uploaded item on the feature stream: it runs beside the previous item's cascade, whose
            coarse-level kernels and fp32-idle tensor layers leave room on the SMs."""
            dimgs, proj, dv, ev = up
            with torch.cuda.stream(feat_stream):
                feat_stream.wait_event(ev)
                dimgs.record_stream(feat_stream)
                feats = self.extract_features(_as_float_images(dimgs))
                fev = torch.cuda.Event()
                fev.record(feat_stream)
            for view in feats:
                for t in view.values():
                    t.record_stream(main)
            return feats, tuple(dimgs.shape[-2:]), fev

        pending = None  # (host dict, event) of the previous item
        it = iter(inputs)
        nxt = next(it, None)
        cur = upload(nxt) if nxt is not None else None
        cur_feats = features_of(cur) if (cur is not None and feat_stream is not None) else None
        while cur is not None:
            nxt = next(it, None)
            dimgs, proj, dv, ev = cur
            if feat_stream is None:
                main.wait_event(ev)
                dimgs.record_stream(main)
                cur = upload(nxt) if nxt is not None else None
                out = self.forward(_as_float_images(dimgs), proj, dv)
            else:
                feats, hw, fev = cur_feats
                # enqueue the next item's upload and FeatureNet first, then this item's cascade: they overlap on the device
                cur = upload(nxt) if nxt is not None else None
                cur_feats = features_of(cur) if cur is not None else None
                main.wait_event(fev)
                out = self.cascade(feats, proj, dv, hw)
                del feats
            done = torch.cuda.Event()
            done.record(main)
            # this item's download goes behind the next upload on the copy stream
            host = {}
            with torch.cuda.stream(side):
                side.wait_event(done)
                finite = torch.empty(1, dtype=torch.bool, pin_memory=True)
                finite.copy_(torch.isfinite(out["depth"]).all().reshape(1), non_blocking=True)
                for k in keys:
                    buf = torch.empty(out[k].shape, dtype=out[k].dtype, pin_memory=True)
                    buf.copy_(out[k], non_blocking=True)
                    out[k].record_stream(side)
                    host[k] = buf
                hev = torch.cuda.Event()
                hev.record(side)
            if pending is not None:
                pending[1].synchronize()
                _check_finite(pending[2])
                yield pending[0]
            pending = (host, hev, finite)
        if pending is not None:
            pending[1].synchronize()
            _check_finite(pending[2])
            yield pending[0]

