"""Ad-hoc (GPU box): hot-path and end-to-end time against the tuning knobs (persistent CTAs per SM of the tensor convs,
H2D view groups of MVSNet.infer)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dmvsnet_b200 import MVSNet, _native, ops, synthetic as syn

H, W, views, nd, ratios = 1184, 1600, 5, [48, 32, 8], [4, 2, 1]
dev = torch.device("cuda:0")
lib = _native.load()
net = MVSNet(nd, ratios, inverse_depth=True)
net.load_state_dict(syn.randomise_regnet_state(net.state_dict(), seed=0))
net = net.to(dev).eval()
imgs_host = syn.make_images(H, W, views, 1, seed=0, natural=True).pin_memory()
proj = syn.make_proj_matrices(H, W, views, 1, num_stages=3)
dv_host = syn.make_depth_values(1, 192, inverse=True)
dv = dv_host.to(dev)
with torch.no_grad():
    feats = net.extract_features(imgs_host.to(dev))


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


def run(label):
    ops.PROFILE = None
    ms = timeit(lambda: net.cascade(feats, proj, dv, (H, W)))
    ops.PROFILE = []
    net.cascade(feats, proj, dv, (H, W)); torch.cuda.synchronize()
    groups = {}
    for tag, a, b, _ in ops.PROFILE:
        groups[tag.split(":")[0]] = groups.get(tag.split(":")[0], 0.0) + a.elapsed_time(b)
    ops.PROFILE = None
    print("%-18s hot path %.2f ms  %s" % (label, ms, {k: round(v, 2) for k, v in groups.items()}), flush=True)


with torch.no_grad():
    for key, values, default in ((b"regnet_streams", (0, 1, 0, 1), 1), (b"tc2_pdl", (0, 1), 1)):
        for v in values:
            lib.dmvs_debug_set(key, v)
            run("%s=%d" % (key.decode(), v))
        lib.dmvs_debug_set(key, default)
    for g in (3,):
        net.infer_view_groups = g
        print("infer_view_groups=%d  e2e %.2f ms" % (g, timeit(lambda: net.infer(imgs_host, proj, dv_host))), flush=True)
