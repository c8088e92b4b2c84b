"""Host-side mirror of the reference's ``networks/module.py`` for the cost-volume hot path.

Same public names, constructor arguments, ``state_dict`` keys and tensor semantics as the
reference, so that checkpoints and callers carry over unchanged; the arithmetic of the regularisation
nets, the warp and the hypothesis sampler runs in libdmvs_b200.so (sm_100a CUDA) through
``dmvsnet_b200.ops`` - and so does ``FeatureNet`` for CUDA tensors; its plain
PyTorch definition remains for CPU tensors and as the ``engine = "cudnn"`` cross-check.

Reference lines each piece answers to are cited per class / function.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

__all__ = ["Conv2d", "Deconv2d", "Conv3d", "Deconv3d", "FeatureNet", "CostRegNet", "CostRegNet_refine", "CostRegNet_part",
           "CostRegNet_part_refine", "homo_warping", "depth_regression", "get_depth_range_samples"]


# --------------------------------------------------------------------------------------------
# conv -> BatchNorm -> ReLU blocks.  reference networks/module.py:28-208.
# Attribute names `conv` / `bn` fix the state_dict keys (`<block>.conv.weight`, `<block>.bn.running_mean` ...).
# --------------------------------------------------------------------------------------------
class _ConvBlock(nn.Module):
    _conv_cls = None
    _bn_cls = None
    _transposed = False

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        assert stride in (1, 2)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.relu = kernel_size, stride, relu
        self.conv = self._conv_cls(in_channels, out_channels, kernel_size, stride=stride, bias=(not bn), **kwargs)
        self.bn = self._bn_cls(out_channels, momentum=bn_momentum) if bn else None

    def forward(self, x):
        y = self.conv(x)
        if self._transposed and y.dim() == 4 and self.stride == 2:
            y = y[:, :, : 2 * x.shape[2], : 2 * x.shape[3]].contiguous()  # module.py:104-106
        if self.bn is not None:
            y = self.bn(y)
        return F.relu(y, inplace=True) if self.relu else y

    def packed(self) -> ops.PackedLayer:
        """Kernel-side parameters (eval-mode BatchNorm as per-channel scale/shift)."""
        bn = None
        if self.bn is not None:
            bn = (self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)
            eps = self.bn.eps
        else:
            eps = 1e-5
        if self.conv.bias is not None:
            raise NotImplementedError("conv blocks with a bias and no BatchNorm are not on the hot path")
        return ops.PackedLayer(self.conv.weight, self._transposed, bn, eps)


class Conv2d(_ConvBlock):
    _conv_cls, _bn_cls = nn.Conv2d, nn.BatchNorm2d


class Deconv2d(_ConvBlock):
    _conv_cls, _bn_cls, _transposed = nn.ConvTranspose2d, nn.BatchNorm2d, True


class Conv3d(_ConvBlock):
    _conv_cls, _bn_cls = nn.Conv3d, nn.BatchNorm3d


class Deconv3d(_ConvBlock):
    _conv_cls, _bn_cls, _transposed = nn.ConvTranspose3d, nn.BatchNorm3d, True


# --------------------------------------------------------------------------------------------
# FeatureNet - reference networks/module.py:274-340.  CUDA tensors run on this library's kernels (fp32 direct
# convolutions + the tcgen05 engine, channel-last outputs for W1); CPU tensors and engine="cudnn" take the plain PyTorch graph.
# --------------------------------------------------------------------------------------------
class FeatureNet(nn.Module):
    def __init__(self, base_channels, num_stage=3, stride=4, mode="fpn", layernorm=False):
        super().__init__()
        assert mode in ("unet", "fpn")
        c = base_channels
        self.mode, self.stride, self.base_channels, self.num_stage, self.layernorm = mode, stride, c, num_stage, layernorm
        self.conv0 = nn.Sequential(Conv2d(3, c, 3, 1, padding=1), Conv2d(c, c, 3, 1, padding=1))
        self.conv1 = nn.Sequential(Conv2d(c, 2 * c, 5, stride=2, padding=2), Conv2d(2 * c, 2 * c, 3, 1, padding=1),
                                   Conv2d(2 * c, 2 * c, 3, 1, padding=1))
        self.conv2 = nn.Sequential(Conv2d(2 * c, 4 * c, 5, stride=2, padding=2), Conv2d(4 * c, 4 * c, 3, 1, padding=1),
                                   Conv2d(4 * c, 4 * c, 3, 1, padding=1))
        # every head emits the main and the `_c` (refine) feature set side by side
        self.out1 = nn.Conv2d(4 * c, 8 * c, 1, bias=False)
        self.inner1 = nn.Conv2d(2 * c, 4 * c, 1, bias=True)
        self.inner2 = nn.Conv2d(c, 4 * c, 1, bias=True)
        self.out2 = nn.Conv2d(4 * c, 4 * c, 3, padding=1, bias=False)
        self.out3 = nn.Conv2d(4 * c, 2 * c, 3, padding=1, bias=False)
        self.out_channels = [4 * c, 2 * c, c]

    # cuDNN may run fp32 convolutions on TF32 tensor cores (torch default); that alone moves the stage-1 cost volume by
    # ~1e-3 relative, i.e. the whole parity budget, so it is off unless the caller opts in.
    allow_tf32 = False
    # "native": this library's fp32 direct-convolution kernels (dmvs_conv2d_f32), features emitted channel-last for the
    # W1 kernels; "cudnn": torch/cuDNN, the reference's own path.  CPU tensors always
    # take the torch path (FeatureNet is above the hot path and keeps a plain PyTorch definition), and so do
    # configurations other than the reference's (fpn, 3 stages).
    engine = "native"
    # native engine: run out2 / out3 on the tensor cores (fp16 hi/lo split operands, fp32 accumulate: same 1e-6 error as the
    # fp32 FMA kernels, 2.6x faster); False keeps them on the fp32 direct convolution
    tensor_heads = True
    tensor_s2 = True  # the two 5x5 stride-2 layers on the tensor engine through a 2x2 pixel-unshuffle (space-to-depth)
    # native engine with tensor heads: the out2 / out3 epilogues also write their maps rounded to fp16 (extra ``stageK_h16`` /
    # ``stageK_c_h16`` entries, ops.HalfFeatures: W1's source-map format).  Set by MVSNet.extract_features around its call;
    # off by default so that FeatureNet.forward returns exactly the reference's keys
    emit_f16 = False

    def forward(self, x):
        if x.is_cuda and self.engine == "native" and not self.training and self.mode == "fpn" and self.num_stage == 3:
            return self._forward_native(x)
        if x.is_cuda:
            with torch.backends.cudnn.flags(enabled=True, benchmark=torch.backends.cudnn.benchmark, allow_tf32=self.allow_tf32):
                return self._forward(x)
        return self._forward(x)

    def _forward(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1(c0)
        c2 = self.conv2(c1)
        out = {}

        def emit(name, t):
            half = t.shape[1] // 2
            out[name], out[name + "_c"] = t.split([half, half], 1)

        emit("stage1", self.out1(c2))
        top = F.interpolate(c2, scale_factor=2, mode="nearest") + self.inner1(c1)
        emit("stage2", self.out2(top))
        top = F.interpolate(top, scale_factor=2, mode="nearest") + self.inner2(c0)
        emit("stage3", self.out3(top))
        return out

    # ------------------------------------------------------------------ native path
    def _state_key(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def __getstate__(self):
        # derived caches (repacked weights) are rebuilt on demand: copy.deepcopy / torch.save / pickle work like on the reference's module
        state = self.__dict__.copy()
        state.pop("_packed", None)
        state.pop("_packed_key", None)
        return state

    def packed(self):
        key = self._state_key()
        if getattr(self, "_packed_key", None) != key:
            def block(m):
                bn = (m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var)
                return ops.PackedConv2d(m.conv.weight, bn=bn, eps=m.bn.eps, stride=m.stride, relu=m.relu)

            def s2d_layer(m):  # 5x5 stride-2 block as a 3x3 stride-1 tensor-core layer on the pixel-unshuffled input
                bn = (m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var)
                return ops.PackedLayer(ops.s2d_weight(m.conv.weight.detach()), False, bn, m.bn.eps)
            pk = {"conv0": [block(m) for m in self.conv0], "conv1": [block(m) for m in self.conv1],
                  "conv2": [block(m) for m in self.conv2],
                  "out1": ops.PackedConv2d(self.out1.weight), "out2": ops.PackedConv2d(self.out2.weight),
                  "out3": ops.PackedConv2d(self.out3.weight),
                  "conv1_s2d": s2d_layer(self.conv1[0]), "conv2_s2d": s2d_layer(self.conv2[0]),
                  "conv1_tc": [self.conv1[1].packed(), self.conv1[2].packed()],
                  "conv2_tc": [self.conv2[1].packed(), self.conv2[2].packed()],
                  "out2_tc": ops.PackedLayer(self.out2.weight, False, None), "out3_tc": ops.PackedLayer(self.out3.weight, False, None),
                  "inner1": ops.PackedConv2d(self.inner1.weight, bias=self.inner1.bias),
                  "inner2": ops.PackedConv2d(self.inner2.weight, bias=self.inner2.bias)}
            self._packed, self._packed_key = pk, key
        return self._packed

    def _forward_native(self, x):
        """Same graph as ``_forward`` on dmvs_conv2d_f32.  Every returned map is a [B,C,h,w] tensor that is physically
        channel-last (what the W1 kernels read in place)."""
        pk = self.packed()
        t = x
        s2d = self.tensor_heads and self.tensor_s2 and x.shape[-1] % 8 == 0 and x.shape[-2] % 4 == 0
        # the two full-resolution layers with few channels (3 -> 8, 8 -> 8) stay on the fp32 direct convolution (the tensor engine
        # measured slower there: 3.57 vs 3.47 ms per DTU view set)
        t = ops.conv2d(t, pk["conv0"][0])
        if s2d:  # conv0.1 hands its output out twice: fp32 NCHW (lateral inner2) and unshuffled cells (conv1.0 on the tensor cores)
            t, cells0 = ops.conv2d(t, pk["conv0"][1], cells=True, s2d=True)
        else:
            t, cells0 = ops.conv2d(t, pk["conv0"][1]), None
        c0 = t
        if self.tensor_heads and t.shape[-1] % 8 == 0:
            # the 3x3 layers behind each stride-2 5x5 (16->16, 32->32; BN + ReLU in the epilogue) run on the tensor cores, and
            # so do the 5x5 stride-2 layers themselves as 3x3 stride-1 layers on the 2x2 pixel-unshuffled input (4x channels);
            # the last 3x3 of a level returns fp32 NCHW for the fp32 consumers (laterals)
            for name in ("conv1", "conv2"):
                if s2d:
                    cells = cells0 if name == "conv1" else ops.s2d_cells(t)
                    cells = ops.conv3d_ch16(cells, pk[name + "_s2d"], relu=True, out_fmt="ch16")
                else:
                    _, cells = ops.conv2d(t, pk[name][0], nchw=False, cells=True)
                cells = ops.conv3d_ch16(cells, pk[name + "_tc"][0], relu=True, out_fmt="ch16")
                t = ops.conv3d_ch16(cells, pk[name + "_tc"][1], relu=True, out_fmt="f32").squeeze(2)
                if name == "conv1":
                    c1 = t
            c2 = t
        else:
            for layer in pk["conv1"]:
                t = ops.conv2d(t, layer)
            c1 = t
            for layer in pk["conv2"]:
                t = ops.conv2d(t, layer)
            c2 = t
        out = {}
        _, out["stage1"], out["stage1_c"] = ops.conv2d(c2, pk["out1"], nchw=False, split_nhwc=True)
        if self.tensor_heads and c1.shape[-1] % 4 == 0:
            # the two 32-channel 3x3 heads (57 % of FeatureNet's flops) on the tcgen05 engine: the laterals emit their sums as
            # fp16 hi/lo cells (top2 only as cells: nobody else reads it), the heads write the channel-last feature sets
            top, cells = ops.conv2d(c1, pk["inner1"], up_add=c2, cells=True)
            if self.emit_f16:  # the heads' epilogues also write W1's fp16 source maps (``<key>_h16``: ops.HalfFeatures)
                out["stage2"], out["stage2_c"], out["stage2_h16"], out["stage2_c_h16"] = ops.conv2d_head_tensor(cells, pk["out2_tc"], True)
            else:
                out["stage2"], out["stage2_c"] = ops.conv2d_head_tensor(cells, pk["out2_tc"])
            _, cells = ops.conv2d(c0, pk["inner2"], up_add=top, nchw=False, cells=True)
            if self.emit_f16:
                out["stage3"], out["stage3_c"], out["stage3_h16"], out["stage3_c_h16"] = ops.conv2d_head_tensor(cells, pk["out3_tc"], True)
            else:
                out["stage3"], out["stage3_c"] = ops.conv2d_head_tensor(cells, pk["out3_tc"])
            return out
        top = ops.conv2d(c1, pk["inner1"], up_add=c2)
        _, out["stage2"], out["stage2_c"] = ops.conv2d(top, pk["out2"], nchw=False, split_nhwc=True)
        top = ops.conv2d(c0, pk["inner2"], up_add=top)
        _, out["stage3"], out["stage3_c"] = ops.conv2d(top, pk["out3"], nchw=False, split_nhwc=True)
        return out


# --------------------------------------------------------------------------------------------
# Regularisation U-Nets - reference networks/module.py:342-436.
# The nn.Modules own the parameters (identical state_dict keys); forward runs in the CUDA library.
# --------------------------------------------------------------------------------------------
_LAYER_ORDER = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv9", "conv11")


class CostRegNet_part(nn.Module):
    _refine = False

    def __init__(self, in_channels, base_channels, stage=0):
        super().__init__()
        c = base_channels
        self.conv0 = Conv3d(in_channels, c, padding=1)
        self.conv1 = Conv3d(c, 2 * c, stride=2, padding=1)
        self.conv2 = Conv3d(2 * c, 2 * c, padding=1)
        self.conv3 = Conv3d(2 * c, 4 * c, stride=2, padding=1)
        self.conv4 = Conv3d(4 * c, 4 * c, padding=1)
        if self._refine:  # 2-D bottleneck once depth has been squeezed to one plane (module.py:411-414)
            self.conv5 = Conv2d(4 * c, 8 * c, 3, stride=2, padding=1)
            self.conv6 = Conv2d(8 * c, 8 * c, 3, padding=1)
            self.conv7 = Deconv2d(8 * c, 4 * c, 3, stride=2, padding=1, output_padding=1)
        else:
            self.conv5 = Conv3d(4 * c, 8 * c, stride=2, padding=1)
            self.conv6 = Conv3d(8 * c, 8 * c, padding=1)
            self.conv7 = Deconv3d(8 * c, 4 * c, stride=2, padding=1, output_padding=1)
        self.conv9 = Deconv3d(4 * c, 2 * c, stride=2, padding=1, output_padding=1)
        self.conv11 = Deconv3d(2 * c, c, stride=2, padding=1, output_padding=1)
        self.prob = nn.Conv3d(c, 2, 3, stride=1, padding=1, bias=False)

    def packed_layers(self) -> List[ops.PackedLayer]:
        layers = [getattr(self, n).packed() for n in _LAYER_ORDER]
        layers.append(ops.PackedLayer(self.prob.weight, False, None))
        return layers

    def forward(self, x, stage=0):
        """Single-branch forward, layer by layer through the CUDA conv kernels. [B,2,D,h,w] -> [B,2,D,h,w]."""
        _require_inference(self)
        L = self.packed_layers()
        c0 = ops.conv3d(x, L[0])
        c2 = ops.conv3d(ops.conv3d(c0, L[1], stride=2), L[2])
        c4 = ops.conv3d(ops.conv3d(c2, L[3], stride=2), L[4])
        y = ops.conv3d(ops.conv3d(c4, L[5], stride=2), L[6])
        y = ops.conv3d(y, L[7], skip=c4)
        y = ops.conv3d(y, L[8], skip=c2)
        y = ops.conv3d(y, L[9], skip=c0)
        return ops.conv3d(y, L[10], relu=False)


class CostRegNet_part_refine(CostRegNet_part):
    _refine = True


def _require_inference(mod: nn.Module) -> None:
    if mod.training:
        raise NotImplementedError(
            "dmvsnet_b200 implements the inference forward of the cost-volume path (eval-mode BatchNorm); "
            "call .eval() first. Training/backward is a later row (SURVEY.md §8f N2).")


class _DualRegNet(nn.Module):
    """cosR_small + cosR_huge over the same cost volume, concatenated to 4 logit channels (module.py:342-357)."""
    _part = CostRegNet_part

    def __init__(self, in_channels, base_channels, stage=0):
        super().__init__()
        if base_channels != 8 or in_channels != 2:
            raise NotImplementedError("the CUDA path is specialised for in_channels=2, base_channels=8 (reference default)")
        self.cosR_small = self._part(in_channels, base_channels, stage=0)
        self.cosR_huge = self._part(in_channels, base_channels, stage=0)
        self._pack: Optional[ops.PackedRegnet] = None
        self._pack_key = None

    def _state_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def __getstate__(self):
        # the cache holds ctypes structs with raw device pointers: not picklable, and stale in a copy anyway
        state = self.__dict__.copy()
        state["_pack"], state["_pack_key"] = None, None
        return state

    def packed(self) -> ops.PackedRegnet:
        """Repacked weights, cached until a parameter/buffer is written, moved or reloaded."""
        key = self._state_key()
        if self._pack is None or key != self._pack_key:
            with torch.no_grad():
                self._pack = ops.PackedRegnet([self.cosR_small.packed_layers(), self.cosR_huge.packed_layers()],
                                              refine=self._part._refine)
            self._pack_key = key
        return self._pack

    def forward(self, x, cost_cells=None, branch_group=None):
        """x: cost volume [B,2,D,h,w] (may be None when ``cost_cells`` - W1's cell-format output - feeds the tensor engine).
        ``branch_group``: a 2-rank process group - rank r runs branch r only and the logit halves are exchanged with one
        all-gather (the two U-Nets are independent, module.py:343-349); every rank must hold the same input."""
        _require_inference(self)
        if branch_group is not None:
            from . import parallel
            return parallel.regnet_branch_sharded(self.packed(), x, cost_cells, branch_group)
        return ops.regnet_forward(self.packed(), x, cost_cells=cost_cells)


class CostRegNet(_DualRegNet):
    _part = CostRegNet_part


class CostRegNet_refine(_DualRegNet):
    _part = CostRegNet_part_refine


# --------------------------------------------------------------------------------------------
# functional seams
# --------------------------------------------------------------------------------------------
def depth_regression(p, depth_values, axis=1):
    """reference networks/module.py:454-460 (kept for API parity; the CUDA heads fuse it)."""
    if depth_values.dim() <= 2:
        depth_values = depth_values.view(*depth_values.shape, 1, 1)
    return torch.sum(p * depth_values, axis=axis)


def get_depth_range_samples(last_depth, ndepth, depth_inteval_pixel, shape=None, next_depth_inteval_pixel=None, inverse=False):
    """reference networks/module.py:556-649.  Returns (samples [B,D,h,w], interval) at the resolution of
    ``last_depth`` (stage 0: ``shape``); ``MVSNet.forward`` uses the fused sample+upsample call instead."""
    if last_depth.dim() == 2:
        return ops.hypotheses_first(last_depth, ndepth, shape, inverse)
    return ops.hypotheses_next(last_depth, ndepth, depth_inteval_pixel, None, inverse)


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """reference networks/module.py:212-251: (warped [B,C,D,H,W], grid [B,D,H,W,2]).

    Kept for API parity only - it materialises exactly the volume the fused W1 kernel exists to avoid,
    and nothing on this repo's forward path calls it.  Implemented with the same torch calls as the reference.
    """
    b, c, h, w = src_fea.shape
    if depth_values.dim() == 2:
        depth_values = depth_values.view(b, -1, 1, 1).expand(-1, -1, h, w)
    d = depth_values.shape[1]
    with torch.no_grad():
        proj = torch.matmul(src_proj, torch.inverse(ref_proj))
        rot, trans = proj[:, :3, :3], proj[:, :3, 3:4]
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=src_fea.device),
                                torch.arange(w, dtype=torch.float32, device=src_fea.device), indexing="ij")
        pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=src_fea.device)))
        pts = torch.matmul(rot, pix.unsqueeze(0).expand(b, -1, -1)).unsqueeze(2) * depth_values.reshape(b, 1, d, -1)
        pts = pts + trans.view(b, 3, 1, 1)
        z = pts[:, 2]
        z = torch.where(z == 0, z + 1e-5, z)
        grid = torch.stack(((pts[:, 0] / z) / ((w - 1) / 2) - 1, (pts[:, 1] / z) / ((h - 1) / 2) - 1), dim=3)
    warped = F.grid_sample(src_fea, grid.view(b, d * h, w, 2), mode="bilinear", padding_mode="zeros", align_corners=True)
    return warped.view(b, c, d, h, w).type(torch.float32), grid.view(b, d, h, w, 2)
