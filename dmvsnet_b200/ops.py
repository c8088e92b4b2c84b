"""Tensor-level wrappers over the C ABI (include/dmvs_b200.h).

Every function takes/returns CUDA fp32 torch tensors, allocates outputs with ``torch.empty`` (so
the caching allocator owns all memory) and enqueues kernels on the current torch stream.
PyTorch is plumbing here: device memory and streams.  There is no fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _native as N


# Optional instrumentation for bench.py: when PROFILE is a list, every op appends (tag, start_event, end_event,
# algorithmic_bytes) recorded on the launching stream; when CAPTURE is a list, warp_corr appends (tag, rt, hyp).
PROFILE = None
PROFILE_ONLY = None  # optional tag prefix (e.g. "w1:"): only those ops are timed - a pair of CUDA events per op costs host time
CAPTURE = None


class _timed:
    def __init__(self, tag: str, nbytes: int = 0):
        self.tag, self.nbytes = tag, nbytes

    def __enter__(self):
        self.on = PROFILE is not None and (PROFILE_ONLY is None or self.tag.startswith(PROFILE_ONLY))
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.on and PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.tag, self.e0, self.e1, self.nbytes))
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("dmvsnet_b200: `%s` must be a CUDA tensor - the cost-volume path has no CPU fallback" % name)
    if t.dtype != torch.float32:
        raise TypeError("dmvsnet_b200: `%s` must be float32, got %s" % (name, t.dtype))
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _device_of(obj):
    if isinstance(obj, torch.Tensor):
        return obj.device if obj.is_cuda else None
    if isinstance(obj, (list, tuple)):
        for o in obj:
            d = _device_of(o)
            if d is not None:
                return d
        return None
    d = getattr(obj, "device", None)  # HalfFeatures
    return d if isinstance(d, torch.device) and d.type == "cuda" else None


def _on_device(fn):
    """Run an op with the device of its first CUDA argument current: the native entry points launch on the CURRENT device and
    on its current stream (``_stream()``), and the > 48 KB shared-memory opt-ins are per device - a model that lives on cuda:1
    while cuda:0 is current would otherwise launch on the wrong device."""
    import functools

    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        dev = _device_of(args) or _device_of(tuple(kwargs.values()))
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return guarded


# ----------------------------------------------------------------------------- projections (K1)
def relative_projections(proj_matrices: torch.Tensor) -> torch.Tensor:
    """[B,N,2,4,4] -> rt [B,N-1,12] (row-major rot 3x3 then trans) on the CPU, fp32.

    Same operations, same order and same library calls as the reference
    (networks/mvsnet.py:133-136 composition, networks/module.py:223-225 ``src @ inverse(ref)``) so the
    homographies are bit-identical to the reference's CPU path.  This is 4x4 host arithmetic on
    N matrices per stage, done once per forward for all stages; the result is uploaded once.
    """
    pm = proj_matrices.detach().to("cpu", torch.float32)
    b, n = pm.shape[0], pm.shape[1]

    def compose(v):
        p = pm[:, v, 0].clone()
        p[:, :3, :4] = torch.matmul(pm[:, v, 1, :3, :3], pm[:, v, 0, :3, :4])
        return p

    ref_inv = torch.inverse(compose(0))
    out = torch.empty(b, n - 1, 12)
    for v in range(1, n):
        m = torch.matmul(compose(v), ref_inv)
        out[:, v - 1, :9] = m[:, :3, :3].reshape(b, 9)
        out[:, v - 1, 9:] = m[:, :3, 3]
    return out


# ----------------------------------------------------------------------------- W1
def _batch_stride(t: torch.Tensor) -> int:
    """Accept channel-sliced views of FeatureNet's output ([B,2C,h,w].split): (c,h,w) dense, any batch stride."""
    b, c, h, w = t.shape
    if t.stride(3) != 1 or t.stride(2) != w or t.stride(1) != h * w:
        return -1
    return t.stride(0) if b > 1 else c * h * w


# which W1 kernel ops.warp_corr uses: "nhwc" = channel-last sources (dmvs_warp_corr_nhwc_f32, sources repacked on the fly
# unless they already are channel-last), "staged" = the same layout, source footprints staged in shared memory by TMA with
# the "nhwc" kernel as the per-tile fallback (dmvs_warp_corr_staged_f32), "nchw" = the original kernel on the reference's
# layout (dmvs_warp_corr_f32)
W1_LAYOUT = "auto"  # "auto": staged where the caller says the hypotheses are spatially coherent, else nhwc
LAST_W1_FLAGS = None  # "staged": the (tile, plane) flags of the most recent call (1 = computed by the fallback pass), for diagnostics


def _nhwc_strides(t: torch.Tensor):
    """(pixel stride, batch stride) if ``t`` [B,C,h,w] is physically channel-last (torch channels_last, possibly a channel
    slice of a wider map), else None."""
    b, c, h, w = t.shape
    ps = t.stride(3)
    if t.stride(1) != 1 or ps < c or ps % 4 or t.stride(2) != w * ps or t.data_ptr() % 16:
        return None
    bs = t.stride(0) if b > 1 else h * w * ps
    return (ps, bs) if bs % 4 == 0 else None


@_on_device
def features_nhwc(t: torch.Tensor) -> torch.Tensor:
    """[B,C,h,w] fp32 (dense (c,h,w), any batch stride) -> the same map physically channel-last, returned as a
    [B,C,h,w] view of a dense [B,h,w,C] buffer (torch channels_last strides), so it can be passed wherever the NCHW
    tensor went."""
    lib = N.load()
    _req(t, "features")
    b, c, h, w = t.shape
    bs = _batch_stride(t)
    if bs < 0:
        t = t.contiguous()
        bs = c * h * w
    y = torch.empty(b, h, w, c, device=t.device, dtype=torch.float32)
    with _timed("w1_layout:C%d_%dx%d" % (c, h, w), 8 * b * c * h * w):
        rc = lib.dmvs_features_nhwc_f32(t.data_ptr(), bs, y.data_ptr(), b, c, h, w, _stream())
    N.check(rc, "dmvs_features_nhwc_f32")
    return y.permute(0, 3, 1, 2)


class HalfFeatures:
    """A feature map rounded to fp16, dense channel-last in memory ([B,h,w,C] halfs): what dmvs_warp_corr_h16_f32 gathers
    from.  ``shape`` is the logical [B,C,h,w] of the fp32 map it stands for."""

    def __init__(self, data: torch.Tensor):
        b, h, w, c = data.shape
        assert data.dtype == torch.float16 and data.stride()[1:] == (w * c, c, 1), "dense [h,w,C] halfs per batch entry"
        self.data = data
        self.shape = torch.Size((b, c, h, w))
        self.device = data.device
        self.bstride = data.stride(0) if b > 1 else h * w * c

    def view_of(self, batch: int, views: int, v: int) -> "HalfFeatures":
        """Entry v of every item of a [batch * views] stack (FeatureNet runs all views in one batched call)."""
        b, h, w, c = self.data.shape
        return HalfFeatures(self.data.view(batch, views, h, w, c)[:, v])

    def record_stream(self, stream) -> None:
        self.data.record_stream(stream)

    def float(self) -> torch.Tensor:
        """Back to an fp32 [B,C,h,w] view-shaped tensor (values already rounded to fp16)."""
        return self.data.float().permute(0, 3, 1, 2)


@_on_device
def features_nhwc_f16(t) -> "HalfFeatures":
    """fp32 [B,C,h,w] (NCHW with any batch stride, or channel-last in memory) -> HalfFeatures (dmvs_features_nhwc_f16)."""
    if isinstance(t, HalfFeatures):
        return t
    lib = N.load()
    _req(t, "features")
    b, c, h, w = t.shape
    cl = _nhwc_strides(t)
    if cl is not None:
        ps, bs = cl
    else:
        ps, bs = 0, _batch_stride(t)
        if bs < 0:
            t = t.contiguous()
            bs = c * h * w
    y = torch.empty(b, h, w, c, device=t.device, dtype=torch.float16)
    with _timed("w1_layout16:C%d_%dx%d" % (c, h, w), 6 * b * c * h * w):
        rc = lib.dmvs_features_nhwc_f16(t.data_ptr(), bs, ps, y.data_ptr(), b, c, h, w, _stream())
    N.check(rc, "dmvs_features_nhwc_f16")
    return HalfFeatures(y)


@_on_device
def warp_corr(features: Sequence[torch.Tensor], rt: torch.Tensor, hyp: torch.Tensor,
              d_range: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None,
              want_f32: bool = True, want_cells: bool = False, layout: Optional[str] = None, coherent: bool = False,
              row0: int = 0):
    """features: N x [B,C,h,w] (reference view first), rt [B,N-1,12] (device), hyp [B,D,h,w] -> cost [B,2,D,h,w].

    ``want_cells``: additionally (or, with ``want_f32=False``, only) emit the cost volume in the cell layout the tensor
    path's first conv reads by TMA (DMVS_FMT_COST2, int32 [B,D,h,w+1,4]); the call then returns ``(cost_or_None, cells)``.
    ``layout`` (default ``W1_LAYOUT``): "nhwc" gathers from channel-last source maps - sources that are not already
    channel-last in memory are repacked first (one read + one write of the map) - "staged" the TMA-staged kernel with the
    "nhwc" kernel as its per-tile fallback, "nchw" the kernel on the reference's layout; "auto" picks "staged" when the caller
    declares the hypotheses spatially ``coherent`` (neighbouring pixels sample neighbouring source positions: sampler planes),
    else "nhwc"."""
    lib = N.load()
    layout = layout or W1_LAYOUT
    if layout == "auto":
        layout = "staged" if coherent else "nhwc"
    if layout not in ("nhwc", "nchw", "staged", "h16"):
        raise ValueError("layout must be 'nhwc', 'nchw', 'staged' or 'h16'")
    if layout == "h16":
        return _warp_corr_h16(features, rt, hyp, d_range, out, want_f32, want_cells, row0)
    if row0:
        raise ValueError("row bands (row0 != 0) need layout='h16'")
    if any(isinstance(f, HalfFeatures) for f in features):
        raise TypeError("fp16 source maps (HalfFeatures) need layout='h16'")
    ref = _req(features[0], "features[0]")
    b, c, h, w = ref.shape
    n_src = len(features) - 1
    if n_src < 1 or n_src > N.MAX_SRC:
        raise ValueError("need 1..%d source views, got %d" % (N.MAX_SRC, n_src))
    for i, f in enumerate(features):
        _req(f, "features[%d]" % i)
        if f.shape != ref.shape:
            raise ValueError("features[%d] has shape %s, expected %s" % (i, tuple(f.shape), tuple(ref.shape)))
    ref_bs, ref_ps = _batch_stride(ref), 0
    if ref_bs < 0 and layout != "nchw" and _nhwc_strides(ref) is not None:
        ref_ps, ref_bs = _nhwc_strides(ref)  # channel-last reference features are read in place as well
    elif ref_bs < 0:
        ref = ref.contiguous()
        ref_bs = c * h * w
    srcs = list(features[1:])
    if layout == "nchw":
        strides = [_batch_stride(f) for f in srcs]
        if min(strides) < 0 or len(set(strides)) != 1:
            srcs = [f.contiguous() for f in srcs]
            strides = [c * h * w] * n_src
        src_ps, src_bs, src_cs = 0, strides[0], 0
    else:
        st = [_nhwc_strides(f) for f in srcs]
        src_cs = 0  # x-corners of a footprint one pixel stride apart
        if any(x is None for x in st) or len(set(st)) != 1:
            srcs = [f if x == (c, h * w * c) else features_nhwc(f) for f, x in zip(srcs, st)]
            st = [(c, h * w * c)] * n_src
        src_ps, src_bs = st[0]
    hyp = _req(hyp, "hyp").contiguous()
    rt = _req(rt, "rt").contiguous()
    d = hyp.shape[1]
    if hyp.shape != (b, d, h, w) or rt.shape != (b, n_src, 12):
        raise ValueError("hyp %s / rt %s do not match features %s" % (tuple(hyp.shape), tuple(rt.shape), tuple(ref.shape)))
    if out is None and want_f32:
        out = torch.empty(b, 2, d, h, w, device=ref.device, dtype=torch.float32)
    cells = torch.empty(b, d, h, w + 1, 4, device=ref.device, dtype=torch.int32) if want_cells else None
    lo, hi = (0, d) if d_range is None else d_range
    flags = torch.empty(lib.dmvs_warp_corr_flag_bytes(b, d, h, w), device=ref.device, dtype=torch.uint8) if layout == "staged" else None
    src_ptrs = (ctypes.c_void_p * n_src)(*[f.data_ptr() for f in srcs])
    if CAPTURE is not None:
        CAPTURE.append(("w1:C%d_D%d_%dx%d" % (c, hi - lo, h, w), rt, hyp))
    # algorithmic bytes: every feature map, the hypotheses and the cost volume cross HBM exactly once
    nbytes = 4 * b * h * w * ((n_src + 1) * c + 3 * (hi - lo))
    with _timed("w1:C%d_D%d_%dx%d" % (c, hi - lo, h, w), nbytes):
        if layout == "nchw":
            rc = lib.dmvs_warp_corr_f32(ref.data_ptr(), ref_bs, src_ptrs, src_bs, n_src, rt.data_ptr(), hyp.data_ptr(),
                                        _ptr(out), _ptr(cells), b, c, d, h, w, lo, hi, _stream())
        elif layout == "nhwc":
            rc = lib.dmvs_warp_corr_nhwc_f32(ref.data_ptr(), ref_bs, ref_ps, src_ptrs, src_bs, src_ps, src_cs, n_src, rt.data_ptr(), hyp.data_ptr(),
                                             _ptr(out), _ptr(cells), b, c, d, h, w, lo, hi, _stream())
        else:
            rc = lib.dmvs_warp_corr_staged_f32(ref.data_ptr(), ref_bs, ref_ps, src_ptrs, src_bs, src_ps, src_cs, n_src, rt.data_ptr(),
                                               hyp.data_ptr(), _ptr(out), _ptr(cells), flags.data_ptr(), b, c, d, h, w, lo, hi, _stream())
    N.check(rc, "dmvs_warp_corr_f32[%s]" % layout)
    if layout == "staged":
        global LAST_W1_FLAGS
        LAST_W1_FLAGS = flags
    return (out, cells) if want_cells else out


def _warp_corr_h16(features, rt, hyp, d_range, out, want_f32, want_cells, row0=0):
    """``warp_corr(layout="h16")``: fp32 reference view, fp16 channel-last sources (fp32 sources are rounded on the fly).
    Row band: the reference map and ``hyp`` may hold only rows [row0, row0 + h) of the view; the sources stay whole."""
    lib = N.load()
    ref = features[0].float() if isinstance(features[0], HalfFeatures) else _req(features[0], "features[0]")
    b, c, h, w = ref.shape
    n_src = len(features) - 1
    if n_src < 1 or n_src > N.MAX_SRC:
        raise ValueError("need 1..%d source views, got %d" % (N.MAX_SRC, n_src))
    srcs = [features_nhwc_f16(f) for f in features[1:]]
    src_rows = srcs[0].shape[2]
    for i, f in enumerate(srcs):
        if f.shape != srcs[0].shape or (f.shape[0], f.shape[1], f.shape[3]) != (b, c, w) or row0 + h > src_rows:
            raise ValueError("features[%d] has shape %s, reference band %s at row %d" % (i + 1, tuple(f.shape), tuple(ref.shape), row0))
    ref_bs, ref_ps = _batch_stride(ref), 0
    if ref_bs < 0 and _nhwc_strides(ref) is not None:
        ref_ps, ref_bs = _nhwc_strides(ref)
    elif ref_bs < 0:
        ref = ref.contiguous()
        ref_bs = c * h * w
    hyp = _req(hyp, "hyp").contiguous()
    rt = _req(rt, "rt").contiguous()
    d = hyp.shape[1]
    if hyp.shape != (b, d, h, w) or rt.shape != (b, n_src, 12):
        raise ValueError("hyp %s / rt %s do not match features %s" % (tuple(hyp.shape), tuple(rt.shape), tuple(ref.shape)))
    if out is None and want_f32:
        out = torch.empty(b, 2, d, h, w, device=ref.device, dtype=torch.float32)
    cells = torch.empty(b, d, h, w + 1, 4, device=ref.device, dtype=torch.int32) if want_cells else None
    lo, hi = (0, d) if d_range is None else d_range
    if len({f.bstride for f in srcs}) != 1:
        srcs = [HalfFeatures(f.data.contiguous()) for f in srcs]
    src_ptrs = (ctypes.c_void_p * n_src)(*[f.data.data_ptr() for f in srcs])
    if CAPTURE is not None:
        CAPTURE.append(("w1:C%d_D%d_%dx%d" % (c, hi - lo, h, w), rt, hyp))
    # algorithmic bytes: SURVEY 8d's figure (fp32 maps once each + hypotheses + cost volume), whatever the storage format
    nbytes = 4 * b * h * w * ((n_src + 1) * c + 3 * (hi - lo))
    with _timed("w1:C%d_D%d_%dx%d" % (c, hi - lo, h, w), nbytes):
        rc = lib.dmvs_warp_corr_h16_f32(ref.data_ptr(), ref_bs, ref_ps, src_ptrs, srcs[0].bstride, c, n_src, rt.data_ptr(), hyp.data_ptr(),
                                        _ptr(out), _ptr(cells), b, c, d, h, w, lo, hi, int(row0), src_rows, _stream())
    N.check(rc, "dmvs_warp_corr_h16_f32")
    return (out, cells) if want_cells else out


@_on_device
def warp_corr_backward(features: Sequence[torch.Tensor], rt: torch.Tensor, hyp: torch.Tensor, grad_cost: torch.Tensor) -> List[torch.Tensor]:
    """Gradients of ``warp_corr``'s cost volume w.r.t. every feature map (reference view first): N x [B,C,h,w] tensors in
    channels_last memory (dmvs_warp_corr_backward_f32).  What autograd would record for networks/mvsnet.py:137-146 and
    module.py:247-249; the sampling grid carries no gradient in the reference (module.py:222), so neither do ``hyp`` / ``rt``."""
    lib = N.load()
    ref = _req(features[0], "features[0]")
    b, c, h, w = ref.shape
    n_src = len(features) - 1
    if n_src < 1 or n_src > N.MAX_SRC:
        raise ValueError("need 1..%d source views, got %d" % (N.MAX_SRC, n_src))
    maps = []
    for i, f in enumerate(features):
        f = _req(f, "features[%d]" % i).detach()
        if f.shape != ref.shape:
            raise ValueError("features[%d] has shape %s, expected %s" % (i, tuple(f.shape), tuple(ref.shape)))
        maps.append(f if _nhwc_strides(f) is not None else features_nhwc(f))
    st = [_nhwc_strides(f) for f in maps[1:]]
    if len(set(st)) != 1:  # the sources share one stride pair in the ABI
        maps[1:] = [f if x == (c, h * w * c) else features_nhwc(f) for f, x in zip(maps[1:], st)]
        st = [(c, h * w * c)] * n_src
    ref_ps, ref_bs = _nhwc_strides(maps[0])
    src_ps, src_bs = st[0]
    hyp = _req(hyp, "hyp").contiguous()
    rt = _req(rt, "rt").contiguous()
    d = hyp.shape[1]
    grad_cost = _req(grad_cost, "grad_cost").contiguous()
    if hyp.shape != (b, d, h, w) or rt.shape != (b, n_src, 12) or grad_cost.shape != (b, 2, d, h, w):
        raise ValueError("hyp %s / rt %s / grad_cost %s do not match features %s"
                         % (tuple(hyp.shape), tuple(rt.shape), tuple(grad_cost.shape), tuple(ref.shape)))
    grads = [torch.empty(b, h, w, c, device=ref.device, dtype=torch.float32) for _ in range(n_src + 1)]
    src_ptrs = (ctypes.c_void_p * n_src)(*[f.data_ptr() for f in maps[1:]])
    gsrc_ptrs = (ctypes.c_void_p * n_src)(*[g.data_ptr() for g in grads[1:]])
    # algorithmic bytes: features and their gradients once each, hypotheses and the cost gradient once
    nbytes = 4 * b * h * w * (2 * (n_src + 1) * c + 3 * d)
    with _timed("w1_bwd:C%d_D%d_%dx%d" % (c, d, h, w), nbytes):
        rc = lib.dmvs_warp_corr_backward_f32(maps[0].data_ptr(), ref_bs, ref_ps, src_ptrs, src_bs, src_ps, n_src, rt.data_ptr(),
                                             hyp.data_ptr(), grad_cost.data_ptr(), grads[0].data_ptr(), gsrc_ptrs, b, c, d, h, w, _stream())
    N.check(rc, "dmvs_warp_corr_backward_f32")
    return [g.permute(0, 3, 1, 2) for g in grads]


class _WarpCorrFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rt, hyp, layout, coherent, *features):
        ctx.save_for_backward(rt, hyp, *features)
        return warp_corr([f.detach() for f in features], rt, hyp, layout=layout, coherent=coherent)

    @staticmethod
    def backward(ctx, grad_cost):
        rt, hyp, *features = ctx.saved_tensors
        grads = warp_corr_backward(features, rt, hyp, grad_cost)
        need = ctx.needs_input_grad[4:]
        return (None, None, None, None) + tuple(g if n else None for g, n in zip(grads, need))


def warp_corr_autograd(features: Sequence[torch.Tensor], rt: torch.Tensor, hyp: torch.Tensor, layout: Optional[str] = None,
                       coherent: bool = False) -> torch.Tensor:
    """Differentiable W1: ``warp_corr`` forward, ``warp_corr_backward`` recorded for autograd (gradients to the feature maps
    only, like the reference: module.py:222 builds the grid under no_grad, mvsnet.py:221 detaches the depth between stages)."""
    return _WarpCorrFunction.apply(rt, hyp.detach(), layout, coherent, *features)


# ----------------------------------------------------------------------------- N1 (FeatureNet)
def _fold_bn(bn, eps: float):
    gamma, beta, mean, var = [t.detach().to(torch.float32) for t in bn]
    scale = gamma / torch.sqrt(var + eps)
    return scale.contiguous(), (beta - mean * scale).contiguous()


class PackedConv2d:
    """One FeatureNet layer for dmvs_conv2d_f32: weights [Cin][K][K][Cout], eval BatchNorm folded to scale/shift (or the
    conv bias as shift)."""

    def __init__(self, weight: torch.Tensor, bn=None, bias: Optional[torch.Tensor] = None, eps: float = 1e-5,
                 stride: int = 1, relu: bool = False):
        w = weight.detach().to(torch.float32)
        self.cout, self.cin, self.k = w.shape[0], w.shape[1], w.shape[2]
        self.w = w.permute(1, 2, 3, 0).contiguous()
        self.scale, self.shift = _fold_bn(bn, eps) if bn is not None else (None, None)
        if bias is not None:
            self.shift = bias.detach().to(torch.float32).contiguous()
        self.stride, self.relu = stride, relu


@_on_device
def conv2d(x: torch.Tensor, layer: PackedConv2d, up_add: Optional[torch.Tensor] = None, nchw: bool = True, split_nhwc: bool = False,
           cells: bool = False, s2d: bool = False):
    """x [B,Cin,H,W] -> conv (+BN/bias, ReLU, + nearest-x2 ``up_add``).  Returns ``y`` ([B,Cout,Ho,Wo] NCHW) when ``nchw``;
    with ``split_nhwc`` additionally the two channel halves as channel-last buffers, each returned as a [B,Cout/2,Ho,Wo]
    view (torch channels_last strides): ``(y_or_None, half0, half1)``; with ``cells`` additionally the
    output as CH16 cells (int32 [B, Cout/4, 1, Ho, Wo, 4]; with ``s2d`` the cells of the 2x2 pixel-unshuffled map,
    int32 [B, Cout, 1, Ho/2, Wo/2, 4]) for the tensor-core layers: ``(y_or_None, cells)``."""
    lib = N.load()
    x = _req(x, "x").contiguous()
    b, cin, hi, wi = x.shape
    if cin != layer.cin:
        raise ValueError("conv2d: input has %d channels, layer expects %d" % (cin, layer.cin))
    pad = layer.k // 2
    ho, wo = (hi + 2 * pad - layer.k) // layer.stride + 1, (wi + 2 * pad - layer.k) // layer.stride + 1
    y = torch.empty(b, layer.cout, ho, wo, device=x.device, dtype=torch.float32) if nchw else None
    h0 = h1 = None
    if split_nhwc:
        h0 = torch.empty(b, ho, wo, layer.cout // 2, device=x.device, dtype=torch.float32)
        h1 = torch.empty_like(h0)
    yc = None
    if cells:
        yc = (torch.empty(b, layer.cout, 1, ho // 2, wo // 2, 4, device=x.device, dtype=torch.int32) if s2d else
              torch.empty(b, layer.cout // 4, 1, ho, wo, 4, device=x.device, dtype=torch.int32))
    if up_add is not None:
        up_add = _req(up_add, "up_add").contiguous()
        if up_add.shape != (b, layer.cout, ho // 2, wo // 2):
            raise ValueError("conv2d: up_add has shape %s, expected %s" % (tuple(up_add.shape), (b, layer.cout, ho // 2, wo // 2)))
    with _timed("featnet:k%ds%d_%dto%d_%dx%d" % (layer.k, layer.stride, cin, layer.cout, ho, wo)):
        rc = lib.dmvs_conv2d_f32(x.data_ptr(), layer.w.data_ptr(), _ptr(layer.scale), _ptr(layer.shift), _ptr(up_add), _ptr(y),
                                 _ptr(h0), _ptr(h1), _ptr(yc), int(bool(cells and s2d)), b, cin, layer.cout, hi, wi, layer.k, layer.stride,
                                 int(layer.relu), _stream())
    N.check(rc, "dmvs_conv2d_f32")
    if split_nhwc:
        return y, h0.permute(0, 3, 1, 2), h1.permute(0, 3, 1, 2)
    if cells:
        return y, yc
    return y


@_on_device
def s2d_cells(x: torch.Tensor) -> torch.Tensor:
    """fp32 [B,C,H,W] -> CH16 cells of the 2x2 pixel-unshuffled map, int32 [B, C, 1, H/2, W/2, 4] (dmvs_features_s2d_cells_f32)."""
    lib = N.load()
    x = _req(x, "x").contiguous()
    b, c, h, w = x.shape
    y = torch.empty(b, c, 1, h // 2, w // 2, 4, device=x.device, dtype=torch.int32)
    with _timed("featnet:s2d_%d_%dx%d" % (c, h, w)):
        rc = lib.dmvs_features_s2d_cells_f32(x.data_ptr(), y.data_ptr(), b, c, h, w, _stream())
    N.check(rc, "dmvs_features_s2d_cells_f32")
    return y


def s2d_weight(w5: torch.Tensor) -> torch.Tensor:
    """5x5 stride-2 (padding 2) conv weight [Cout,C,5,5] -> the equivalent 3x3 stride-1 (padding 1) weight [Cout,4C,3,3] on the
    2x2 pixel-unshuffled input (channel (dy*2+dx)*C + c): output (y,x) reads input rows 2y-2..2y+2 = blocks y-1..y+1, so the
    kernel is zero-padded to 6x6 (index 5 is never read) and regrouped."""
    cout, c = w5.shape[0], w5.shape[1]
    w6 = torch.zeros(cout, c, 6, 6, dtype=w5.dtype, device=w5.device)
    w6[:, :, :5, :5] = w5
    w6 = w6.reshape(cout, c, 3, 2, 3, 2)               # [co][c][bh][dy][bw][dx]
    return w6.permute(0, 3, 5, 1, 2, 4).reshape(cout, 4 * c, 3, 3).contiguous()


@_on_device
def conv2d_head_tensor(cells: torch.Tensor, layer: "PackedLayer", want_f16: bool = False):
    """FeatureNet's bare 3x3 heads out2 / out3 (module.py:326-336, Cin = 32, no BN / ReLU / bias) on the tcgen05 engine:
    CH16 cells [B, 8, 1, H, W, 4] -> the two feature sets as channel-last maps, returned as [B,Cout/2,H,W] views.
    ``want_f16``: the epilogue also writes both sets rounded to fp16 (DMVS_FMT_NHWC2_F16) - W1's source-map format - and the call
    returns ``(set0, set1, HalfFeatures0, HalfFeatures1)``."""
    lib = N.load()
    b, planes, d, h, w, _ = cells.shape
    if planes != 8 or d != 1 or layer.cin != 32 or layer.kd != 1 or layer.w_tc is None:
        raise ValueError("conv2d_head_tensor: expects 32-channel cells and a packed 2-D 3x3 layer")
    half = layer.cout // 2
    n = 2 * b * h * w * half
    buf = torch.empty(n + (n // 2 if want_f16 else 0), device=cells.device, dtype=torch.float32)
    y = buf[:n].view(2, b, h, w, half)
    cl = layer.c_struct()
    with _timed("featnet:tc3x3_32to%d_%dx%d" % (layer.cout, h, w)):
        rc = lib.dmvs_conv3d_ch16(cells.data_ptr(), 0, ctypes.byref(cl), None, buf.data_ptr(), b, 32, layer.cout, 1, h, w, 1, 1, 0, 0,
                                  N.FMT_NHWC2_F16 if want_f16 else N.FMT_NHWC2, _stream())
    N.check(rc, "dmvs_conv3d_ch16")
    if want_f16:
        y16 = buf[n:].view(torch.float16).view(2, b, h, w, half)
        return y[0].permute(0, 3, 1, 2), y[1].permute(0, 3, 1, 2), HalfFeatures(y16[0]), HalfFeatures(y16[1])
    return y[0].permute(0, 3, 1, 2), y[1].permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------- R1
# Range of the tensor engine: activations between layers, cost-volume cells and weights are carried as fp16 hi + fp16 lo
# (value = hi + lo).  |x| <= 65504 keeps hi finite; beyond it hi = inf, lo = -inf and the MMA produces NaN, which propagates to the
# depth map - MVSNet.infer / infer_many check the finiteness of what they return and raise.  Feature maps, BN-folded activations and
# costs of a trained DMVSNet are O(1) - O(100); a network whose activations leave the range must run with DEFAULT_ENGINE = "fp32".
FP16_SPLIT_MAX = 65504.0


class PackedLayer:
    """Device-side parameters of one conv block in the layout the kernels read."""

    def __init__(self, weight: torch.Tensor, transposed: bool, bn: Optional[Tuple[torch.Tensor, ...]], eps: float = 1e-5):
        w = weight.detach().to(torch.float32)
        if w.dim() == 4:  # 2-D conv of the refine bottleneck -> kd = 1
            w = w.unsqueeze(2)
        # Conv: [Cout,Cin,kd,kh,kw]; ConvTranspose: [Cin,Cout,kd,kh,kw]  ->  [tap][Cin][Cout]
        w = w.permute(2, 3, 4, 0, 1) if transposed else w.permute(2, 3, 4, 1, 0)
        taps = w.shape[0] * w.shape[1] * w.shape[2]
        cin, cout = w.shape[3], w.shape[4]
        w = w.reshape(taps, cin, cout)
        cout_w = (cout + 3) // 4 * 4
        if cout_w != cout:
            w = torch.nn.functional.pad(w, (0, cout_w - cout))
        self.w = w.contiguous()
        # the tensor engine splits every operand into fp16 hi + lo: a weight beyond the fp16 range has hi = inf (and NaN products)
        if w.numel() and not bool((w.abs() <= FP16_SPLIT_MAX).all()):
            raise ValueError("conv weight beyond +-%.0f cannot be split into fp16 hi/lo for the tensor engine (use ops.DEFAULT_ENGINE = 'fp32')" % FP16_SPLIT_MAX)
        if cin % 8 == 0:
            self.w_tc = _pack_tensor_core(w[:, :, :cout])
        elif cin == 2 and taps == 27 and not transposed:
            self.w_tc = _pack_tensor_core_kw(w[:, :, :cout])
        else:
            self.w_tc = None
        self.w_tc_kd = None
        self.w_tc_kw = None
        if cin == 8 and cout == 2 and taps == 27 and not transposed:
            self.w_tc_kd = _pack_tensor_core_prob(w[:, :, :cout])
            self.w_tc_kw = _pack_tensor_core_prob_wide(w[:, :, :cout])
        elif cin == 16 and cout == 8 and taps == 27 and transposed:
            self.w_tc_kd = _pack_tensor_core_tr_fold(w[:, :, :cout])
        elif cin == 16 and cout == 16 and taps == 27 and not transposed:   # conv2
            self.w_tc_kd = _pack_tensor_core_kf(self.w_tc)
        elif cin == 32 and cout == 32 and taps == 27 and not transposed:   # conv4
            self.w_tc_kd = _pack_tensor_core_kf_split(w[:, :, :cout])
        elif cin == 2 and cout == 8 and taps == 27 and not transposed:     # conv0
            self.w_tc_kd = _pack_tensor_core_kf(self.w_tc, kh_only=True)
        self.kd = 3 if taps == 27 else 1
        self.cin, self.cout = cin, cout
        self.transposed = transposed
        if bn is not None:
            gamma, beta, mean, var = [t.detach().to(torch.float32) for t in bn]
            self.scale = (gamma / torch.sqrt(var + eps)).contiguous()
            self.shift = (beta - mean * self.scale).contiguous()
        else:
            self.scale = self.shift = None

    def c_struct(self) -> N.ConvLayer:
        return N.ConvLayer(self.w.data_ptr(), _ptr(self.scale), _ptr(self.shift), _ptr(self.w_tc), _ptr(self.w_tc_kd), _ptr(self.w_tc_kw))


def _pack_tensor_core(w: torch.Tensor) -> torch.Tensor:
    """[tap][Cin][Cout] fp32 -> the fp16 hi/lo image of include/dmvs_b200.h (dmvs_conv_layer.w_tc):
    [chunk j][tap][kc][n = 2*Cout_p][8 halfs]; columns [0,Cout_p) hold hi(W) for both K halves (they multiply A_hi and
    A_lo), columns [Cout_p, 2*Cout_p) hold lo(W) for the A_hi half only."""
    taps, cin, cout = w.shape
    cout_p = max(8, (cout + 7) // 8 * 8)
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    img = torch.zeros(cin // 8, taps, 2, 2 * cout_p, 8, dtype=torch.float16, device=w.device)
    hi_r = hi.reshape(taps, cin // 8, 8, cout).permute(1, 0, 3, 2)  # [j][tap][n][k]
    lo_r = lo.reshape(taps, cin // 8, 8, cout).permute(1, 0, 3, 2)
    img[:, :, 0, :cout] = hi_r
    img[:, :, 1, :cout] = hi_r
    img[:, :, 0, cout_p:cout_p + cout] = lo_r
    return img.contiguous()


def _pack_tensor_core_prob(w: torch.Tensor) -> torch.Tensor:
    """`prob` (8 -> 2): depth tap folded into N.  [27][8][2] -> [1][9 (kh,kw)][kc][n = 16][8 halfs]; column n = 4*kd + co holds
    hi(W) (both K halves), n = 4*kd + 2 + co holds lo(W) (A_hi half only); columns 12..15 stay zero."""
    taps, cin, cout = w.shape
    assert taps == 27 and cin == 8 and cout == 2
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    hi = hi.reshape(3, 9, 8, 2)  # [kd][(kh,kw)][k][co]
    lo = lo.reshape(3, 9, 8, 2)
    img = torch.zeros(1, 9, 2, 16, 8, dtype=torch.float16, device=w.device)
    for kd in range(3):
        for co in range(2):
            img[0, :, 0, 4 * kd + co] = hi[kd, :, :, co]
            img[0, :, 1, 4 * kd + co] = hi[kd, :, :, co]
            img[0, :, 0, 4 * kd + 2 + co] = lo[kd, :, :, co]
    return img.contiguous()


def _pack_tensor_core_prob_wide(w: torch.Tensor) -> torch.Tensor:
    """`prob` (8 -> 2) for the wide-tile kernel: depth tap and kw folded into N.  [27][8][2] -> [1][3 (kh)][kc][n = 48][8 halfs];
    column n = 16*kd + 4*kw + co holds hi(W) (both K halves), n = 16*kd + 4*kw + 2 + co holds lo(W) (A_hi half only)."""
    taps, cin, cout = w.shape
    assert taps == 27 and cin == 8 and cout == 2
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    hi = hi.reshape(3, 3, 3, 8, 2)  # [kd][kh][kw][k][co]
    lo = lo.reshape(3, 3, 3, 8, 2)
    img = torch.zeros(1, 3, 2, 48, 8, dtype=torch.float16, device=w.device)
    for kd in range(3):
        for kw in range(3):
            for co in range(2):
                n = 16 * kd + 4 * kw + co
                img[0, :, 0, n] = hi[kd, :, kw, :, co]
                img[0, :, 1, n] = hi[kd, :, kw, :, co]
                img[0, :, 0, n + 2] = lo[kd, :, kw, :, co]
    return img.contiguous()


def _pack_tensor_core_kf(per_tap: torch.Tensor, kh_only: bool = False) -> torch.Tensor:
    """Depth tap folded into N for the z-marching kernels (csrc/conv_kf.cu).  `per_tap` is the image of ``_pack_tensor_core``
    ([j][27 taps (kd,kh,kw)][kc][nb][8]) or, with ``kh_only``, of ``_pack_tensor_core_kw`` ([1][9 taps (kd,kh)][kc][nb][8]);
    the result is [j][taps / 3][kc][n = 3*nb][8] with the column blocks ordered kd = 0, 1, 2."""
    j, taps, kc, nb, e = per_tap.shape
    assert taps == (9 if kh_only else 27)
    t = per_tap.reshape(j, 3, taps // 3, kc, nb, e).permute(0, 2, 3, 1, 4, 5)
    return t.reshape(j, taps // 3, kc, 3 * nb, e).contiguous()


def _pack_tensor_core_kf_split(w: torch.Tensor) -> torch.Tensor:
    """Depth tap folded into N with the lo(W) product in the K dimension (csrc/conv_kf.cu, kind SW: conv4, 32 -> 32).
    [27][Cin][Cout] -> one flat buffer holding two images with columns n = kd * Cout + co:
      hi image [chunk j = Cin/8][9 taps (kh,kw)][kc = 2][n][8 halfs]: hi(W[8j+k]) for kc = 0 AND 1 (the A_hi and A_lo halves of chunk j);
      lo image [pair i = Cin/16][9 taps][kc = 2][n][8 halfs]: lo(W[16i+k]) for kc = 0, lo(W[16i+8+k]) for kc = 1 (the A_hi halves of the
      pair's two chunks)."""
    taps, cin, cout = w.shape
    assert taps == 27 and cin % 16 == 0 and cout % 8 == 0
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    cj = cin // 8
    hi_r = hi.reshape(3, 9, cj, 8, cout).permute(2, 1, 0, 4, 3)       # [j][tap9][kd][co][k]
    lo_r = lo.reshape(3, 9, cj // 2, 2, 8, cout).permute(2, 1, 3, 0, 5, 4)  # [pair][tap9][kc][kd][co][k]
    img_hi = torch.empty(cj, 9, 2, 3 * cout, 8, dtype=torch.float16, device=w.device)
    img_hi[:, :, 0] = hi_r.reshape(cj, 9, 3 * cout, 8)
    img_hi[:, :, 1] = hi_r.reshape(cj, 9, 3 * cout, 8)
    img_lo = lo_r.reshape(cj // 2, 9, 2, 3 * cout, 8)
    return torch.cat([img_hi.reshape(-1), img_lo.reshape(-1)]).contiguous()


def _pack_tensor_core_tr_fold(w: torch.Tensor) -> torch.Tensor:
    """ConvTranspose3d(k3, s2, p1, op1) with the taps folded by input shift.  [27][Cin][Cout] -> [chunk j][shift s = (sz,sy,sx)]
    [kc][n = 8 classes x 2*Cout_p][8 halfs].  Output parity class c = (pz,py,px) along an axis: even outputs (p = 0) only see
    tap k = 1 of the unshifted input; odd outputs (p = 1) see tap k = 2 of the unshifted and tap k = 0 of the +1-shifted input.
    The MMA of shift s therefore carries, in class block c, the weights of tap (k_z,k_y,k_x) with k = 1 (p = 0, s = 0),
    k = 2 (p = 1, s = 0), k = 0 (p = 1, s = 1), and zeros where p = 0 meets s = 1.  Inside a class block the columns are
    those of ``_pack_tensor_core``: [hi | lo]."""
    taps, cin, cout = w.shape
    assert taps == 27 and cin % 8 == 0
    cout_p = max(8, (cout + 7) // 8 * 8)
    nb = 2 * cout_p
    per_tap = _pack_tensor_core(w)  # [j][27][kc][nb][8]
    img = torch.zeros(cin // 8, 8, 2, 8 * nb, 8, dtype=torch.float16, device=w.device)
    for s in range(8):
        sh = ((s >> 2) & 1, (s >> 1) & 1, s & 1)
        for c in range(8):
            par = ((c >> 2) & 1, (c >> 1) & 1, c & 1)
            k = []
            for p_, s_ in zip(par, sh):
                if p_ == 0 and s_ == 1:
                    k = None
                    break
                k.append(1 if p_ == 0 else (0 if s_ == 1 else 2))
            if k is None:
                continue
            tap = (k[0] * 3 + k[1]) * 3 + k[2]
            img[:, s, :, c * nb:(c + 1) * nb] = per_tap[:, tap]
    return img.contiguous()


def _pack_tensor_core_kw(w: torch.Tensor) -> torch.Tensor:
    """conv0 (Cin = 2): K packed along kw.  [27][2][Cout] -> [1][9 taps (kd,kh)][kc][n][8 halfs]; the 16 K entries of a tap
    are 4 voxels (x-1, x | x+1, x+2) x [hi c0, hi c1, lo c0, lo c1]; voxel x+2 is outside the 3-wide kernel -> zeros."""
    taps, cin, cout = w.shape
    assert taps == 27 and cin == 2
    cout_p = max(8, (cout + 7) // 8 * 8)
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    hi = hi.reshape(9, 3, 2, cout)  # [(kd,kh)][kw][c][n]
    lo = lo.reshape(9, 3, 2, cout)
    img = torch.zeros(1, 9, 2, 2 * cout_p, 8, dtype=torch.float16, device=w.device)
    for v in range(3):                      # voxel slot v <-> kw = v
        kc, base = divmod(v, 2)
        base *= 4
        for c in range(2):
            img[0, :, kc, :cout, base + c] = hi[:, v, c]                    # A_hi * W_hi
            img[0, :, kc, :cout, base + 2 + c] = hi[:, v, c]                # A_lo * W_hi
            img[0, :, kc, cout_p:cout_p + cout, base + c] = lo[:, v, c]     # A_hi * W_lo
    return img.contiguous()


# engine used by conv3d / regnet_forward unless the caller says otherwise: "tensor" = tcgen05 split-fp16 where a
# specialisation exists (fp32 kernels elsewhere), "fp32" = CUDA-core fp32 everywhere.
DEFAULT_ENGINE = "tensor"
PAIR_CONV0 = True  # run conv0 of both regularisation branches as one launch (N = 32); False = one launch per branch
_ENGINES = {"fp32": N.ENGINE_FP32, "tensor": N.ENGINE_TENSOR}


@_on_device
def conv3d(x: torch.Tensor, layer: PackedLayer, stride: int = 1, relu: bool = True,
           skip: Optional[torch.Tensor] = None, engine: Optional[str] = None) -> torch.Tensor:
    """One conv block on a [B,Cin,D,H,W] tensor (kd = 1 layers take D as a batch of planes)."""
    lib = N.load()
    x = _req(x, "x").contiguous()
    b, cin, di, hi, wi = x.shape
    if cin != layer.cin:
        raise ValueError("Cin mismatch: input %d, layer %d" % (cin, layer.cin))
    if layer.transposed:
        do, ho, wo = (2 * di if layer.kd == 3 else di), 2 * hi, 2 * wi
    elif stride == 2:
        do, ho, wo = ((di - 1) // 2 + 1 if layer.kd == 3 else di), (hi - 1) // 2 + 1, (wi - 1) // 2 + 1
    else:
        do, ho, wo = di, hi, wi
    y = torch.empty(b, layer.cout, do, ho, wo, device=x.device, dtype=torch.float32)
    if skip is not None:
        skip = _req(skip, "skip").contiguous()
        if skip.shape != y.shape:
            raise ValueError("skip shape %s != output shape %s" % (tuple(skip.shape), tuple(y.shape)))
    cl = layer.c_struct()
    rc = lib.dmvs_conv3d_f32(x.data_ptr(), ctypes.byref(cl), _ptr(skip), y.data_ptr(), b, cin, layer.cout, di, hi, wi,
                             layer.kd, 2 if layer.transposed else stride, int(layer.transposed), int(relu),
                             _ENGINES[engine or DEFAULT_ENGINE], _stream())
    N.check(rc, "dmvs_conv3d_f32")
    return y


_FMT = {"f32": N.FMT_F32, "ch16": N.FMT_CH16, "ch16p": N.FMT_CH16P}


@_on_device
def to_ch16(x: torch.Tensor, parity_split: bool = False) -> torch.Tensor:
    """fp32 [B,C,D,H,W] -> the tensor path's cell layout (include/dmvs_b200.h DMVS_FMT_CH16 / CH16P), as a flat byte-equal
    fp32-sized buffer viewed as int32 [B, C/4 planes, D, H, W, 4]."""
    lib = N.load()
    x = _req(x, "x").contiguous()
    b, c, d, h, w = x.shape
    y = torch.empty(b, c // 4, d, h, w, 4, device=x.device, dtype=torch.int32)
    rc = lib.dmvs_convert_layout(x.data_ptr(), y.data_ptr(), b, c, d, h, w, N.FMT_CH16P if parity_split else N.FMT_CH16, 1, _stream())
    N.check(rc, "dmvs_convert_layout")
    return y


@_on_device
def from_ch16(y: torch.Tensor, channels: int, parity_split: bool = False) -> torch.Tensor:
    lib = N.load()
    b, planes, d, h, w, _ = y.shape
    x = torch.empty(b, channels, d, h, w, device=y.device, dtype=torch.float32)
    rc = lib.dmvs_convert_layout(y.data_ptr(), x.data_ptr(), b, channels, d, h, w, N.FMT_CH16P if parity_split else N.FMT_CH16, 0, _stream())
    N.check(rc, "dmvs_convert_layout")
    return x


@_on_device
def conv3d_ch16(x: torch.Tensor, layer: PackedLayer, stride: int = 1, relu: bool = True, skip: Optional[torch.Tensor] = None,
                out_fmt: str = "ch16", in_cells: bool = False) -> torch.Tensor:
    """One block of the tensor path.  x: cells from ``to_ch16`` (parity split for stride 2) or fp32 [B,2,D,H,W] for conv0;
    skip: parity-split cells; returns cells (or fp32 when out_fmt == "f32")."""
    lib = N.load()
    if layer.cin == 2 and in_cells:  # cost cells from warp_corr(want_cells=True): int32 [B, D, H, W+1, 4]
        b, di, hi, wi = x.shape[0], x.shape[1], x.shape[2], x.shape[3] - 1
    elif layer.cin == 2:
        x = _req(x, "x").contiguous()
        b, _, di, hi, wi = x.shape
    else:
        b, _, di, hi, wi, _ = x.shape
    if layer.transposed:
        do, ho, wo = (2 * di if layer.kd == 3 else di), 2 * hi, 2 * wi
    elif stride == 2:
        do, ho, wo = ((di - 1) // 2 + 1 if layer.kd == 3 else di), (hi - 1) // 2 + 1, (wi - 1) // 2 + 1
    else:
        do, ho, wo = di, hi, wi
    if out_fmt == "f32":
        y = torch.empty(b, layer.cout, do, ho, wo, device=x.device, dtype=torch.float32)
    else:
        y = torch.empty(b, layer.cout // 4, do, ho, wo, 4, device=x.device, dtype=torch.int32)
    cl = layer.c_struct()
    rc = lib.dmvs_conv3d_ch16(x.data_ptr(), int(in_cells), ctypes.byref(cl), _ptr(skip), y.data_ptr(), b, layer.cin, layer.cout, di, hi, wi,
                              layer.kd, 2 if layer.transposed else stride, int(layer.transposed), int(relu), _FMT[out_fmt], _stream())
    N.check(rc, "dmvs_conv3d_ch16")
    return y


class PackedRegnet:
    """Both branches of a CostRegNet / CostRegNet_refine, repacked for dmvs_regnet_forward_f32."""

    def __init__(self, branches: Sequence[Sequence[PackedLayer]], refine: bool):
        assert len(branches) == 2 and all(len(b) == N.REGNET_LAYERS for b in branches)
        self.layers = branches  # keeps the device tensors alive
        self.refine = refine
        self.c_branches = (N.RegnetBranch * 2)()
        for i, br in enumerate(branches):
            for j, layer in enumerate(br):
                self.c_branches[i].layer[j] = layer.c_struct()
        # conv0 of both branches as one 2 -> 16 layer (they read the same cost volume; include/dmvs_b200.h: conv0_pair)
        a, b = branches[0][0], branches[1][0]
        self.pair = None
        if (PAIR_CONV0 and a.cin == 2 and a.cout == 8 and b.cin == 2 and b.cout == 8 and a.kd == 3 and not a.transposed
                and a.scale is not None and b.scale is not None):
            w = torch.cat([a.w[:, :, :8], b.w[:, :, :8]], 2)
            img = _pack_tensor_core_kw(w)
            self.pair = (img, torch.cat([a.scale, b.scale]).contiguous(), torch.cat([a.shift, b.shift]).contiguous(),
                         _pack_tensor_core_kf(img, kh_only=True))
            self.c_branches[0].conv0_pair = N.ConvLayer(None, self.pair[1].data_ptr(), self.pair[2].data_ptr(), self.pair[0].data_ptr(),
                                                        self.pair[3].data_ptr(), None)


@_on_device
def regnet_forward(pack: PackedRegnet, cost: Optional[torch.Tensor], engine: Optional[str] = None,
                   cost_cells: Optional[torch.Tensor] = None, branch_mask: int = 3, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cost [B,2,D,h,w] (and / or its cell form from ``warp_corr(want_cells=True)``) -> logits [B,4,D,h,w].
    ``branch_mask`` (bit 0 = cosR_small -> channels 0,1; bit 1 = cosR_huge -> channels 2,3) runs a subset of the two independent
    branches and writes only their channels of ``out`` (multi-GPU branch sharding, parallel.regnet_branch_sharded)."""
    lib = N.load()
    if cost is not None:
        cost = _req(cost, "cost").contiguous()
        b, c, d, h, w = cost.shape
        if c != 2:
            raise ValueError("cost volume must have 2 channels, got %d" % c)
        dev = cost.device
    else:
        if cost_cells is None or (engine or DEFAULT_ENGINE) != "tensor":
            raise ValueError("regnet_forward needs the fp32 cost volume unless the tensor engine gets cost_cells")
        b, d, h, w = cost_cells.shape[0], cost_cells.shape[1], cost_cells.shape[2], cost_cells.shape[3] - 1
        dev = cost_cells.device
    logits = out if out is not None else torch.empty(b, 4, d, h, w, device=dev, dtype=torch.float32)
    if logits.shape != (b, 4, d, h, w) or not logits.is_contiguous():
        raise ValueError("regnet_forward: `out` must be a contiguous [B,4,D,h,w] tensor")
    nbytes = lib.dmvs_regnet_workspace_bytes(int(pack.refine), b, d, h, w)
    ws = torch.empty((nbytes + 3) // 4, device=dev, dtype=torch.float32)
    with _timed("regnet%s:D%d_%dx%d" % ("_refine" if pack.refine else "", d, h, w), 4 * b * 6 * d * h * w):
        rc = lib.dmvs_regnet_forward_branches_f32(pack.c_branches, int(pack.refine), _ptr(cost), _ptr(cost_cells), logits.data_ptr(),
                                                  ws.data_ptr(), ws.numel() * 4, b, d, h, w, _ENGINES[engine or DEFAULT_ENGINE],
                                                  int(branch_mask), _stream())
    N.check(rc, "dmvs_regnet_forward_branches_f32")
    return logits


# ----------------------------------------------------------------------------- E1 / E2
def _scalar(v, device) -> torch.Tensor:
    if isinstance(v, torch.Tensor):
        return v.detach().to(device=device, dtype=torch.float32).reshape(1)
    return torch.full((1,), float(v), device=device, dtype=torch.float32)


@_on_device
def depth_head(logits: torch.Tensor, hyp: torch.Tensor, interval, want_prob: bool = True):
    lib = N.load()
    logits = _req(logits, "logits").contiguous()
    hyp = _req(hyp, "hyp").contiguous()
    b, c, d, h, w = logits.shape
    if c != 4 or hyp.shape != (b, d, h, w):
        raise ValueError("logits %s / hyp %s mismatch" % (tuple(logits.shape), tuple(hyp.shape)))
    dev = logits.device
    prob = torch.empty_like(logits) if want_prob else None
    d4 = torch.empty(b, 4, h, w, device=dev)
    hyp_c = torch.empty(b, 4, h, w, device=dev)
    conf = torch.empty(b, h, w, device=dev)
    iv = _scalar(interval, dev)
    with _timed("head:D%d_%dx%d" % (d, h, w), 4 * b * h * w * ((9 if want_prob else 5) * d + 9)):
        rc = lib.dmvs_depth_head_f32(logits.data_ptr(), hyp.data_ptr(), iv.data_ptr(), _ptr(prob), d4.data_ptr(), hyp_c.data_ptr(),
                                     conf.data_ptr(), b, d, h, w, _stream())
    N.check(rc, "dmvs_depth_head_f32")
    return prob, d4, hyp_c, conf


@_on_device
def refine_head(logits_c: torch.Tensor, hyp_c: torch.Tensor, interval, alpha: float = 5.0):
    lib = N.load()
    logits_c = _req(logits_c, "logits_c").contiguous()
    hyp_c = _req(hyp_c, "hyp_c").contiguous()
    b, c, d, h, w = logits_c.shape
    if c != 4 or d != 4 or hyp_c.shape != (b, 4, h, w):
        raise ValueError("refine head expects logits [B,4,4,h,w] and hypotheses [B,4,h,w]")
    dev = logits_c.device
    depth = torch.empty(b, h, w, device=dev)
    conf = torch.empty(b, h, w, device=dev)
    d4 = torch.empty(b, 4, h, w, device=dev)
    iv = _scalar(interval, dev)
    with _timed("head_refine:%dx%d" % (h, w), 4 * b * h * w * 26):
        rc = lib.dmvs_refine_head_f32(logits_c.data_ptr(), hyp_c.data_ptr(), iv.data_ptr(), float(alpha), depth.data_ptr(),
                                      conf.data_ptr(), d4.data_ptr(), b, h, w, _stream())
    N.check(rc, "dmvs_refine_head_f32")
    return depth, conf, d4


# ----------------------------------------------------------------------------- S1
@_on_device
def hypotheses_first(depth_values: torch.Tensor, ndepth: int, shape: Sequence[int], inverse: bool):
    lib = N.load()
    dv = _req(depth_values, "depth_values").contiguous()
    b, nd = dv.shape
    h, w = int(shape[0]), int(shape[1])
    hyp = torch.empty(b, ndepth, h, w, device=dv.device)
    interval = torch.empty((), device=dv.device)
    with _timed("sampler:D%d_%dx%d" % (ndepth, h, w), 4 * b * ndepth * h * w):
        rc = lib.dmvs_hypotheses_first_f32(dv.data_ptr(), nd, hyp.data_ptr(), interval.data_ptr(), b, ndepth, h, w, int(inverse),
                                           _stream())
    N.check(rc, "dmvs_hypotheses_first_f32")
    return hyp, interval


@_on_device
def hypotheses_next(last_depth: torch.Tensor, ndepth: int, interval_pixel, shape: Optional[Sequence[int]], inverse: bool):
    """Per-pixel checkerboard ranges around ``last_depth`` [B,h0,w0], upsampled to ``shape`` (None: no upsample)."""
    lib = N.load()
    ld = _req(last_depth, "last_depth").contiguous()
    b, h0, w0 = ld.shape
    h, w = (h0, w0) if shape is None else (int(shape[0]), int(shape[1]))
    hyp = torch.empty(b, ndepth, h, w, device=ld.device)
    interval = torch.empty((), device=ld.device)
    ip = _scalar(interval_pixel, ld.device)
    with _timed("sampler:D%d_%dx%d" % (ndepth, h, w), 4 * b * (ndepth * h * w + h0 * w0)):
        rc = lib.dmvs_hypotheses_next_f32(ld.data_ptr(), ip.data_ptr(), hyp.data_ptr(), interval.data_ptr(), b, ndepth, h0, w0, h, w,
                                          int(inverse), _stream())
    N.check(rc, "dmvs_hypotheses_next_f32")
    return hyp, interval
