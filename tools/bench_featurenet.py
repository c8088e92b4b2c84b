"""Ad-hoc: FeatureNet (torch/cuDNN) timing variants on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmvsnet_b200 import MVSNet, synthetic as syn
torch.backends.cudnn.benchmark = True
net = MVSNet([48, 32, 8], [4, 2, 1]).cuda().eval()
imgs = syn.make_images(1184, 1600, 5, 1, natural=True).cuda()

def timeit(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

with torch.no_grad():
    for tf32 in (False, True):
        net.feature.allow_tf32 = tf32
        print("tf32", tf32, "per-view loop       %.2f ms" % timeit(lambda: [net.feature(imgs[:, v]) for v in range(5)]))
        print("tf32", tf32, "batched 5 views     %.2f ms" % timeit(lambda: net.feature(imgs[0])))
        x_cl = imgs[0].contiguous(memory_format=torch.channels_last)
        net_cl = net.feature.to(memory_format=torch.channels_last)
        print("tf32", tf32, "batched channels_last %.2f ms" % timeit(lambda: net_cl(x_cl)))
        net.feature.to(memory_format=torch.contiguous_format)
    # half precision for reference
    f16 = MVSNet([48, 32, 8], [4, 2, 1]).cuda().eval().feature.half()
    xh = imgs[0].half()
    print("fp16 batched %.2f ms" % timeit(lambda: f16(xh)))
