"""Ad-hoc (GPU box): MVSNet.infer_many throughput with and without FeatureNet(k+1) running beside cascade(k)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dmvsnet_b200 import MVSNet, synthetic as syn

H, W, views, nd, ratios = 1184, 1600, 5, [48, 32, 8], [4, 2, 1]
dev = torch.device("cuda:0")
net = MVSNet(nd, ratios, inverse_depth=True)
net.load_state_dict(syn.randomise_regnet_state(net.state_dict(), seed=0))
net = net.to(dev).eval()
imgs = syn.make_images(H, W, views, 1, seed=0, natural=True).pin_memory()
proj = syn.make_proj_matrices(H, W, views, 1, num_stages=3)
dv = syn.make_depth_values(1, 192, inverse=True)
ref = None
for rep in range(2):
    for overlap, conc in ((False, 1), (True, 1), (True, 2), (True, 3)):
        net.overlap_features, net.concurrent_items = overlap, conc
        for n, timed in ((4, False), (12, True)):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            outs = list(net.infer_many([(imgs, proj, dv)] * n))
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            if timed:
                print("overlap_features=%s concurrent_items=%d  %.2f ms per item over %d items" % (overlap, conc, 1e3 * dt / n, n), flush=True)
        if ref is None:
            ref = outs[-1]["depth"].clone()
        assert all(torch.equal(o["depth"], ref) for o in outs)
