"""Ad-hoc (GPU box): FeatureNet's 32-channel 3x3 heads on the tcgen05 engine (kd = 1) vs the fp32 direct convolution."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dmvsnet_b200 import ops

def timeit(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

g = torch.Generator().manual_seed(0)
for (cout, h, w) in ((16, 1184, 1600), (32, 592, 800), (16, 64, 96)):
    x = torch.randn(5, 32, h, w, generator=g).cuda()
    wt = (torch.randn(cout, 32, 3, 3, generator=g) * 0.06).cuda()
    layer = ops.PackedLayer(wt, False, None)
    cells = ops.to_ch16(x.unsqueeze(2))
    y = ops.conv3d_ch16(cells, layer, relu=False, out_fmt="f32").squeeze(2)
    torch.backends.cudnn.allow_tf32 = False
    want = F.conv2d(x.double(), wt.double(), padding=1).float()
    err = float((y - want).abs().max() / want.abs().max())
    pk = ops.PackedConv2d(wt)
    y2 = ops.conv2d(x, pk)
    err2 = float((y2 - want).abs().max() / want.abs().max())
    print("32->%d %dx%d  tensor %.3f ms (err %.2e)   to_ch16 %.3f ms   fp32 direct %.3f ms (err %.2e)" % (
        cout, h, w, timeit(lambda: ops.conv3d_ch16(cells, layer, relu=False, out_fmt="f32")), err,
        timeit(lambda: ops.to_ch16(x.unsqueeze(2))), timeit(lambda: ops.conv2d(x, pk)), err2), flush=True)
