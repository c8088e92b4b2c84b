"""Host-side logic and the C-ABI surface, without a GPU."""
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dmvs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dmvs_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(native_lib):
    from dmvsnet_b200 import _native
    declared = _declared_symbols()
    assert len(declared) >= 11
    for name in declared:
        assert hasattr(native_lib, name), "libdmvs_b200.so does not export %s" % name
    assert sorted(_native.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert native_lib.dmvs_abi_version() == _native.ABI_VERSION == 20
    assert native_lib.dmvs_launch_count() >= 0  # a process-wide counter: other tests of the same session may have launched already


def test_library_is_sm100a_only(native_lib):
    import subprocess
    from dmvsnet_b200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_bad_arguments_return_errors_not_crashes(native_lib):
    # argument validation happens before any CUDA call, so it is testable without a device
    rc = native_lib.dmvs_warp_corr_f32(None, 0, None, 0, 1, None, None, None, None, 1, 32, 4, 8, 8, 0, 4, None)
    assert rc == -2 and b"null" in native_lib.dmvs_last_error()
    rc = native_lib.dmvs_warp_corr_nhwc_f32(None, 0, 0, None, 0, 8, 0, 1, None, None, None, None, 1, 8, 4, 8, 8, 0, 4, None)
    assert rc == -2 and b"null" in native_lib.dmvs_last_error()
    rc = native_lib.dmvs_warp_corr_staged_f32(None, 0, 0, None, 0, 8, 0, 1, None, None, None, None, None, 1, 8, 4, 8, 8, 0, 4, None)
    assert rc == -2 and b"null" in native_lib.dmvs_last_error()
    assert native_lib.dmvs_warp_corr_flag_bytes(2, 48, 296, 400) == 2 * 37 * 13 * 48
    rc = native_lib.dmvs_warp_corr_backward_f32(None, 0, 8, None, 0, 8, 1, None, None, None, None, None, 1, 8, 4, 8, 8, None)
    assert rc == -2 and b"null" in native_lib.dmvs_last_error()
    rc = native_lib.dmvs_geo_consistency_dynamic_f32(None, None, None, 3, 8, 8, 0.25, 1 / 1300, None, None, None, None, None, None, None)
    assert rc == -2 and b"null" in native_lib.dmvs_last_error()
    rc = native_lib.dmvs_backproject_world_f32(None, None, 8, 8, None, None)
    assert rc == -2 and b"null" in native_lib.dmvs_last_error()
    rc = native_lib.dmvs_features_nhwc_f32(None, 0, None, 1, 8, 8, 8, None)
    assert rc == -2 and b"null" in native_lib.dmvs_last_error()
    assert native_lib.dmvs_regnet_workspace_bytes(0, 1, 8, 16, 16) > 0
    assert native_lib.dmvs_regnet_workspace_bytes(0, 0, 8, 16, 16) == 0


def test_workspace_formula(native_lib):
    # main net, B=1, D=8, 16x16: levels (8,16,16) (4,8,8) (2,4,4) (1,2,2)
    v = [8 * 16 * 16, 4 * 8 * 8, 2 * 4 * 4, 1 * 2 * 2]
    # level 0 holds conv0 of BOTH branches (16 channels, conv0_pair); everything else (conv11's output, two buffers per coarser
    # level) exists once per branch so that the branches can run on two streams
    want = 4 * (2 * 8 * v[0] + 2 * (8 * v[0] + 2 * 16 * v[1] + 2 * 32 * v[2] + 2 * 64 * v[3]))
    assert native_lib.dmvs_regnet_workspace_bytes(0, 1, 8, 16, 16) == want
    # refine net: D 4 -> 2 -> 1, then a 2-D level
    v = [4 * 16 * 16, 2 * 8 * 8, 1 * 4 * 4, 1 * 2 * 2]
    want = 4 * (2 * 8 * v[0] + 2 * (8 * v[0] + 2 * 16 * v[1] + 2 * 32 * v[2] + 2 * 64 * v[3]))
    assert native_lib.dmvs_regnet_workspace_bytes(1, 1, 4, 16, 16) == want


def test_state_dict_contract():
    from dmvsnet_b200 import MVSNet
    net = MVSNet([48, 32, 8], [4, 2, 1], inverse_depth=True)
    sd = net.state_dict()
    assert len(sd) == 787
    assert sum(p.numel() for p in net.parameters()) == 2673048
    assert tuple(sd["cost_regularization.0.cosR_small.conv0.conv.weight"].shape) == (8, 2, 3, 3, 3)
    assert tuple(sd["cost_regularization_refine.2.cosR_huge.conv5.conv.weight"].shape) == (64, 32, 3, 3)
    assert tuple(sd["cost_regularization.1.cosR_huge.prob.weight"].shape) == (2, 8, 3, 3, 3)
    assert tuple(sd["cost_regularization.2.cosR_small.conv7.conv.weight"].shape) == (64, 32, 3, 3, 3)
    assert "cost_regularization.0.cosR_small.conv0.bn.num_batches_tracked" in sd
    assert tuple(sd["feature.out1.weight"].shape) == (64, 32, 1, 1)


@pytest.mark.reference
def test_state_dict_identical_to_reference():
    import contextlib
    import io
    import sys
    sys.path.insert(0, "/root/reference")
    with contextlib.redirect_stdout(io.StringIO()):
        from networks import mvsnet as MV
        ref = MV.MVSNet([48, 32, 8], [4, 2, 1])
    from dmvsnet_b200 import MVSNet
    mine = MVSNet([48, 32, 8], [4, 2, 1])
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    mine.load_state_dict(a)  # strict
    x = torch.rand(1, 3, 64, 96)
    ref.eval(), mine.eval()
    with torch.no_grad():
        fa, fb = ref.feature(x), mine.feature(x)
    assert all(torch.equal(fa[k], fb[k]) for k in fa)


def test_weight_packing_and_cache():
    from dmvsnet_b200 import MVSNet, synthetic as syn
    net = MVSNet([8, 8, 8], [4, 2, 1])
    net.load_state_dict(syn.randomise_regnet_state(net.state_dict()))
    net.eval()
    reg = net.cost_regularization[1]
    pk = reg.packed()
    assert reg.packed() is pk  # cached
    br = reg.cosR_huge
    # conv2: [Cout=16,Cin=16,3,3,3] -> [27][16][16]
    w = br.conv2.conv.weight
    assert torch.equal(pk.layers[1][2].w[5, 3, 7], w[7, 3, 0, 1, 2])
    # conv7 is transposed: [Cin=64,Cout=32,3,3,3] -> [27][64][32]
    w = br.conv7.conv.weight
    assert torch.equal(pk.layers[1][7].w[26, 60, 30], w[60, 30, 2, 2, 2])
    # prob: Cout=2 padded to 4
    assert tuple(pk.layers[1][10].w.shape) == (27, 8, 4) and float(pk.layers[1][10].w[:, :, 2:].abs().max()) == 0
    bn = br.conv2.bn
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    assert torch.allclose(pk.layers[1][2].scale, scale) and torch.allclose(pk.layers[1][2].shift, bn.bias - bn.running_mean * scale)
    # writing a parameter invalidates the cache (load_state_dict copies in place)
    net.load_state_dict(syn.randomise_regnet_state(net.state_dict(), seed=3))
    assert reg.packed() is not pk
    rf = net.cost_regularization_refine[0].packed()
    assert rf.refine and rf.layers[0][5].kd == 1 and rf.layers[0][7].kd == 1 and rf.layers[0][7].transposed


def test_cpu_tensors_are_rejected_loudly():
    from dmvsnet_b200 import MVSNet, ops
    net = MVSNet([8], [4]).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.rand(1, 2, 3, 32, 32), {"stage1": torch.zeros(1, 2, 2, 4, 4)}, torch.rand(1, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.warp_corr([torch.zeros(1, 8, 8, 8)] * 2, torch.zeros(1, 1, 12), torch.zeros(1, 2, 8, 8))
    with pytest.raises(NotImplementedError, match="eval"):
        MVSNet([8], [4]).train().cost_regularization[0](torch.zeros(1, 2, 8, 8, 8))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dmvsnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("dmvs_oracle_unused", ""), "%s mentions the oracle" % f


def test_shim_resolves_reference_import_path():
    import importlib
    import sys
    shim = os.path.join(ROOT, "dmvsnet_b200", "shim")
    sys.path.insert(0, shim)
    for m in [k for k in sys.modules if k == "networks" or k.startswith("networks.")]:
        del sys.modules[m]
    try:
        mv = importlib.import_module("networks.mvsnet")
        from dmvsnet_b200 import MVSNet
        assert mv.MVSNet is MVSNet and hasattr(mv, "homo_warping") and hasattr(mv, "get_depth_range_samples")
        md = importlib.import_module("networks.module")
        assert hasattr(md, "CostRegNet") and hasattr(md, "FeatureNet") and hasattr(md, "CostRegNet_part_refine")
    finally:
        sys.path.remove(shim)
        for m in [k for k in sys.modules if k == "networks" or k.startswith("networks.")]:
            del sys.modules[m]


def test_synthetic_rig_is_in_frustum():
    """the look-at rig keeps most plane-sweep samples inside the source images."""
    from dmvsnet_b200 import synthetic as syn
    from oracle import dmvs_oracle as O
    h, w = 74, 100  # DTU stage-1 grid / 4
    proj = syn.make_proj_matrices(h * 4, w * 4, 5)["stage1"]
    hyp, _ = O.depth_hypotheses(syn.make_depth_values(), 48, None, (h, w))
    ref_p = O.compose_projection(proj[:, 0])
    fr = []
    for v in range(1, 5):
        rot, tr = O.relative_projection(O.compose_projection(proj[:, v]), ref_p)
        g = O.sampling_grid(rot, tr, hyp)
        fr.append(float(((g.abs() <= 1).all(-1)).float().mean()))
    assert min(fr) > 0.8, fr


def test_shift_folded_transposed_conv_packing_reproduces_conv_transpose3d():
    """conv11's tensor-core weights with the 27 taps folded by input shift (ops._pack_tensor_core_tr_fold): emulate the 8 MMAs
    per channel chunk (A = the input shifted by (sz,sy,sx), N = 8 parity-class blocks of [hi | lo] columns) in torch and
    compare with F.conv_transpose3d - this pins the class/shift/tap bookkeeping without a GPU."""
    import torch.nn.functional as F
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(7)
    cin, cout = 16, 8
    w = 0.2 * torch.randn(cin, cout, 3, 3, 3, generator=g)
    x = torch.randn(1, cin, 3, 4, 5, generator=g)
    want = F.conv_transpose3d(x, w, stride=2, padding=1, output_padding=1)
    layer = ops.PackedLayer(w, True, None)
    img = layer.w_tc_kd.float()
    assert img.shape == (2, 8, 2, 128, 8)
    d, h, wd = x.shape[2:]
    xpad = F.pad(x, (0, 1, 0, 1, 0, 1))
    got = torch.zeros_like(want)
    for c in range(8):
        pz, py, px = (c >> 2) & 1, (c >> 1) & 1, c & 1
        acc = torch.zeros(cout, d, h, wd)
        for s in range(8):
            sz, sy, sx = (s >> 2) & 1, (s >> 1) & 1, s & 1
            for j in range(2):
                blk = img[j, s, :, c * 16:(c + 1) * 16]                 # [kc][16 columns][8 k]
                assert torch.equal(blk[1, :8], blk[0, :8]) and float(blk[1, 8:].abs().max()) == 0.0  # A_lo only meets W_hi
                weff = blk[0, :8] + blk[0, 8:]                           # hi + lo, [co][k]
                xs = xpad[0, 8 * j:8 * j + 8, sz:sz + d, sy:sy + h, sx:sx + wd]
                acc += torch.einsum("ok,kdhw->odhw", weff, xs)
        got[0, :, pz::2, py::2, px::2] = acc
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5


def test_conv0_pair_packing():
    """conv0 of both regularisation branches as one 2 -> 16 layer: K-packed weights with Cout = 16, BN vectors concatenated."""
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(8)

    def branch():
        layers = [ops.PackedLayer(torch.randn(8, 2, 3, 3, 3, generator=g), False,
                                  (torch.rand(8, generator=g) + 0.5, torch.randn(8, generator=g), torch.randn(8, generator=g), torch.rand(8, generator=g) + 0.5))]
        layers += [ops.PackedLayer(torch.randn(8, 8, 3, 3, 3, generator=g), False, None) for _ in range(10)]
        return layers
    a, b = branch(), branch()
    pack = ops.PackedRegnet([a, b], refine=False)
    assert pack.pair is not None and pack.pair[0].shape == (1, 9, 2, 32, 8)
    assert torch.equal(pack.pair[0][:, :, :, :8], a[0].w_tc[:, :, :, :8]) and torch.equal(pack.pair[0][:, :, :, 8:16], b[0].w_tc[:, :, :, :8])
    assert torch.equal(pack.pair[0][:, :, :, 16:24], a[0].w_tc[:, :, :, 8:16]) and torch.equal(pack.pair[0][:, :, :, 24:32], b[0].w_tc[:, :, :, 8:16])
    assert torch.equal(pack.pair[1], torch.cat([a[0].scale, b[0].scale]))


def test_channel_last_stride_detection():
    from dmvsnet_b200 import ops
    t = torch.zeros(2, 16, 6, 8).contiguous(memory_format=torch.channels_last)
    assert ops._nhwc_strides(t) == (16, 6 * 8 * 16)
    assert ops._nhwc_strides(t.split([8, 8], 1)[1]) == (16, 6 * 8 * 16)
    assert ops._nhwc_strides(torch.zeros(2, 16, 6, 8)) is None
    assert ops._batch_stride(torch.zeros(2, 32, 6, 8).split([16, 16], 1)[1]) == 32 * 48
    assert ops._batch_stride(t) == -1


def test_space_to_depth_weight_transform():
    """FeatureNet's 5x5 stride-2 layers run as 3x3 stride-1 layers on the 2x2 pixel-unshuffled input: ops.s2d_weight must make the
    two convolutions identical (channel order (dy*2+dx)*C + c, as dmvs_features_s2d_cells_f32 writes it)."""
    import torch.nn.functional as F
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 12, 16, generator=g, dtype=torch.float64)
    w = torch.randn(16, 8, 5, 5, generator=g, dtype=torch.float64)
    want = F.conv2d(x, w, stride=2, padding=2)
    xs = F.pixel_unshuffle(x, 2)  # torch orders the channels c*4 + dy*2 + dx
    b, _, h, wd = xs.shape
    xs = xs.reshape(b, 8, 4, h, wd).permute(0, 2, 1, 3, 4).reshape(b, 32, h, wd)
    got = F.conv2d(xs, ops.s2d_weight(w), stride=1, padding=1)
    assert float((got - want).abs().max()) < 1e-10


def test_fusion_source_matrices_layout():
    """fusion.source_matrices: the [60] block the consistency kernel reads = the six matrices of filter/pcd.py:164-191."""
    from dmvsnet_b200 import fusion, synthetic as syn
    proj = syn.make_proj_matrices(64, 96, 3, 1, num_stages=3)["stage3"]
    k0, e0, k1, e1 = proj[0, 0, 1, :3, :3], proj[0, 0, 0], proj[0, 2, 1, :3, :3], proj[0, 2, 0]
    m = fusion.source_matrices(k0, e0, k1.numpy(), e1.numpy())     # numpy and torch inputs both accepted
    assert m.shape == (60,) and m.dtype == torch.float32
    t1 = e1 @ torch.linalg.inv(e0)
    t2 = e0 @ torch.linalg.inv(e1)
    assert torch.equal(m[:9].view(3, 3), torch.linalg.inv(k0)) and torch.equal(m[9:21].view(3, 4), t1[:3])
    assert torch.equal(m[21:30].view(3, 3), k1) and torch.equal(m[30:39].view(3, 3), torch.linalg.inv(k1))
    assert torch.equal(m[39:51].view(3, 4), t2[:3]) and torch.equal(m[51:60].view(3, 3), k0)
    # a point of the reference view at depth d maps into the source view and back onto itself
    x = torch.tensor([40.0, 25.0, 1.0]) * 650.0
    p_src = t1[:3, :3] @ (torch.linalg.inv(k0) @ x) + t1[:3, 3]
    back = t2[:3, :3] @ p_src + t2[:3, 3]
    assert float(((k0 @ back) / (k0 @ back)[2] - torch.tensor([40.0, 25.0, 1.0])).abs().max()) < 1e-2


def test_fusion_matrices_numpy_and_torch_forms_agree():
    """The dynamic-threshold kernel takes the matrices numpy computes (dypcd_tanks.py:66-91), the fixed-threshold kernel the ones
    torch computes (pcd.py:164-191): same 60-float block layout, values equal to fp32 rounding of two LAPACK paths."""
    import numpy as np
    import torch
    from dmvsnet_b200 import fusion, synthetic as syn
    proj = syn.make_proj_matrices(48, 64, 3, 1, num_stages=3)["stage3"]
    ks, es = proj[0, :, 1, :3, :3].float(), proj[0, :, 0].float()
    a = fusion.source_matrices(ks[0], es[0], ks[1], es[1])
    b = fusion.source_matrices_numpy(ks[0].numpy(), es[0].numpy(), ks[1].numpy(), es[1].numpy())
    assert a.shape == b.shape == (60,) and b.dtype == torch.float32
    assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max())
    inv_k = np.linalg.inv(ks[0].numpy())
    assert np.array_equal(b[:9].numpy().reshape(3, 3), inv_k)                      # block 0: inv(K_ref), numpy's own bits
    assert np.array_equal(b[21:30].numpy().reshape(3, 3), ks[1].numpy())           # block 2: K_src
    assert np.array_equal(b[51:60].numpy().reshape(3, 3), ks[0].numpy())           # block 5: K_ref


def test_module_copies_and_pickles_like_the_reference():
    """derived caches (repacked weights as ctypes structs, CUDA streams) must not leak into copy.deepcopy / pickle."""
    import copy
    import io
    import pickle
    import torch
    from dmvsnet_b200 import MVSNet
    net = MVSNet([8, 8, 8], [4, 2, 1]).eval()
    reg = net.cost_regularization[0]
    reg._pack, reg._pack_key = object(), ("stale",)          # stand-ins for the ctypes cache a forward leaves behind
    net.feature._packed, net.feature._packed_key = {"x": lambda: 0}, ("stale",)
    twin = copy.deepcopy(net)
    assert twin.cost_regularization[0]._pack is None and not hasattr(twin.feature, "_packed")
    blob = pickle.dumps(net)
    assert sorted(pickle.loads(blob).state_dict()) == sorted(net.state_dict())
    buf = io.BytesIO()
    torch.save(net, buf)


def _emulate_folded(x, img, cout, cout_p, lo_in_k=False):
    """Numerical model of csrc/conv_kf.cu on the CPU: x [Cin, D, H, W] fp32; `img` a folded weight image.  Every input plane s gets
    the accumulator P[s][n] (n = kd * nb + column); the output plane t is P[t-1][kd=0] + P[t][kd=1] + P[t+1][kd=2]."""
    import torch.nn.functional as F
    cin, d, h, w = x.shape
    hi = x.to(torch.float16).float()
    lo = (x - hi).to(torch.float16).float()
    xp_hi, xp_lo = F.pad(hi, (1, 1, 1, 1)), F.pad(lo, (1, 1, 1, 1))
    cj = cin // 8
    if lo_in_k:   # kind SW: [hi image (cj chunks) | lo image (cj / 2 chunk pairs)], columns [kd][cout]
        nb = cout
        n_hi = cj * 9 * 2 * 3 * nb * 8
        img_hi = img[:n_hi].float().view(cj, 9, 2, 3 * nb, 8)
        img_lo = img[n_hi:].float().view(cj // 2, 9, 2, 3 * nb, 8)
    else:
        nb = 2 * cout_p
        img_hi = img.float().view(cj, 9, 2, 3 * nb, 8)
    acc = torch.zeros(d, 3 * nb, h, w)
    for tap in range(9):
        kh, kw = divmod(tap, 3)
        a_hi = xp_hi[:, :, kh:kh + h, kw:kw + w]
        a_lo = xp_lo[:, :, kh:kh + h, kw:kw + w]
        for j in range(cj):
            acc += torch.einsum("nk,kdhw->dnhw", img_hi[j, tap, 0], a_hi[8 * j:8 * j + 8])
            acc += torch.einsum("nk,kdhw->dnhw", img_hi[j, tap, 1], a_lo[8 * j:8 * j + 8])
        if lo_in_k:
            for i in range(cj // 2):
                acc += torch.einsum("nk,kdhw->dnhw", img_lo[i, tap, 0], a_hi[16 * i:16 * i + 8])
                acc += torch.einsum("nk,kdhw->dnhw", img_lo[i, tap, 1], a_hi[16 * i + 8:16 * i + 16])
    out = torch.zeros(cout, d, h, w)
    for t in range(d):
        for kd in range(3):
            s = t + kd - 1
            if 0 <= s < d:
                blk = acc[s, kd * nb:(kd + 1) * nb]
                out[:, t] += blk[:cout] if lo_in_k else blk[:cout] + blk[cout_p:cout_p + cout]
    return out


@pytest.mark.parametrize("cin,cout,split", [(16, 16, False), (32, 32, True)])
def test_folded_weight_images_reproduce_the_convolution(cin, cout, split):
    """The depth-tap-folded weight images of ops.PackedLayer (conv2: [hi|lo] column blocks per kd; conv4: lo(W) in the K dimension)
    under a CPU model of the kernel's accumulate-per-input-plane scheme == F.conv3d, to the 2^-22 of the dropped lo*lo product."""
    import torch.nn.functional as F
    from dmvsnet_b200 import ops
    g = torch.Generator().manual_seed(cin)
    wgt = torch.randn(cout, cin, 3, 3, 3, generator=g) * 0.1
    x = torch.randn(cin, 4, 6, 7, generator=g)
    w = wgt.permute(2, 3, 4, 1, 0).reshape(27, cin, cout).contiguous()
    if split:
        img = ops._pack_tensor_core_kf_split(w)
    else:
        img = ops._pack_tensor_core_kf(ops._pack_tensor_core(w))
    got = _emulate_folded(x, img.reshape(-1) if split else img, cout, max(8, (cout + 7) // 8 * 8), lo_in_k=split)
    want = F.conv3d(x[None], wgt, None, padding=1)[0]
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 2e-6, err
