"""GPU box: where do the depth-tap-folded kernels lose their time?  `prob` / conv0 at the DTU stage-2 grid with the epilogue or the
MMAs switched off (dmvs_debug_set("kf_dbg")), CUDA-event time per launch."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from dmvsnet_b200 import _native, ops
lib = _native.load()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
d, h, w = 32, 592, 800
if len(sys.argv) > 1:
    d, h, w = [int(v) for v in sys.argv[1].split("x")]
print("prob volume %d x %d x %d (%.0f MB of cells), conv2 at half of it" % (d, h, w, d * h * w * 32 / 1e6), flush=True)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


wp = torch.randn(2, 8, 3, 3, 3, generator=g) * 0.1
prob = ops.PackedLayer(wp.to(dev), False, None)
x8 = ops.to_ch16(torch.randn(1, 8, d, h, w, device=dev))
w2 = torch.randn(16, 16, 3, 3, 3, generator=g) * 0.05
bn = tuple(t.to(dev) for t in (torch.ones(16), torch.zeros(16), torch.zeros(16), torch.ones(16)))
conv2 = ops.PackedLayer(w2.to(dev), False, bn)
x16 = ops.to_ch16(torch.randn(1, 16, d // 2, h // 2, w // 2, device=dev))
for kf, mw, npr, wide in ((0, 0, 0, 0), (1, 2, 1, 0), (1, 1, 1, 0), (1, 2, 1, 1)):  # one TMA thread; at most two issuing threads
    for dbg in (0, 1, 2, 3):
        if kf == 0 and dbg:
            continue
        lib.dmvs_debug_set(b"kf", kf); lib.dmvs_debug_set(b"kf_mw", mw); lib.dmvs_debug_set(b"kf_dbg", dbg); lib.dmvs_debug_set(b"kf_wide", wide)
        tp = timeit(lambda: ops.conv3d_ch16(x8, prob, relu=False, out_fmt="f32"))
        t2 = timeit(lambda: ops.conv3d_ch16(x16, conv2, relu=True, out_fmt="ch16p")) if kf else float("nan")
        print("wide=%d kf=%d issuers=%d producers=%d dbg=%d (1: no epilogue work, 2: no MMAs)  prob %.1f us   conv2 %.1f us" % (wide, kf, mw, npr, dbg, tp, t2), flush=True)
lib.dmvs_debug_set(b"kf", 1); lib.dmvs_debug_set(b"kf_mw", 0); lib.dmvs_debug_set(b"kf_dbg", 0); lib.dmvs_debug_set(b"kf_wide", 0)
