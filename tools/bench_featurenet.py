"""Ad-hoc: FeatureNet timing on the GPU box - this library's fp32 direct-conv engine vs torch/cuDNN variants, per layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmvsnet_b200 import MVSNet, synthetic as syn, ops
torch.backends.cudnn.benchmark = True
net = MVSNet([48, 32, 8], [4, 2, 1]).cuda().eval()
imgs = syn.make_images(1184, 1600, 5, 1, natural=True).cuda()


def timeit(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    net.feature.engine = "native"
    print("native fp32 direct conv, batched 5 views %.2f ms" % timeit(lambda: net.feature(imgs[0])))
    ops.PROFILE = []
    net.feature(imgs[0]); torch.cuda.synchronize()
    for tag, e0, e1, _ in ops.PROFILE:
        print("   %-40s %.3f ms" % (tag, e0.elapsed_time(e1)))
    ops.PROFILE = None
    net.feature.engine = "cudnn"
    if "--cudnn" in sys.argv:
        for tf32 in (False, True):
            net.feature.allow_tf32 = tf32
            print("cudnn tf32", tf32, "batched 5 views     %.2f ms" % timeit(lambda: net.feature(imgs[0])))
