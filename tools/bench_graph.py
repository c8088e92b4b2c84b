"""Ad-hoc (GPU box): the cascade (and the whole forward) replayed from a CUDA graph against the eager launch sequence.

    python tools/bench_graph.py [--reps 20]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dmvsnet_b200 import MVSNet, _native, ops, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
H, W, views, nd, ratios = 1184, 1600, 5, [48, 32, 8], [4, 2, 1]
dev = torch.device("cuda:0")
lib = _native.load()
net = MVSNet(nd, ratios, inverse_depth=True)
net.load_state_dict(syn.ridge_regnet_state(net.state_dict(), seed=0))
net = net.to(dev).eval()
net.DepthNet.return_prob_volume = True
proj = syn.make_proj_matrices(H, W, views, 1, num_stages=3)
dv = syn.make_depth_values(1, 192, inverse=True).to(dev)
imgs = syn.make_scene_images(H, W, views, proj["stage3"], seed=0).to(dev)
rts = [ops.relative_projections(proj["stage%d" % (s + 1)]).to(dev) for s in range(3)]


def timeit(fn, n):
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
    feats = net.add_half_features([{k: v.to(dev) for k, v in f.items()} for f in syn.make_scene_features(H, W, views, proj, seed=0)])
    run_c = lambda: net.cascade(feats, proj, dv, (H, W), rts=rts)
    run_f = lambda: net.cascade(net.extract_features(imgs), proj, dv, (H, W), rts=rts)
    import time
    run_k = lambda: net.cascade(feats, proj, dv, (H, W))
    for name, fn in (("cascade, K1 on the host per call", run_k), ("cascade, rts given", run_c)):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            fn()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print("%s: host enqueue %.3f ms per call, wall %.3f ms per call, CUDA events %.3f ms" % (name, (t1 - t0) * 100, (t2 - t0) * 100, timeit(fn, args.reps)), flush=True)
    t0 = time.perf_counter()
    for _ in range(10):
        [ops.relative_projections(proj["stage%d" % (s + 1)]) for s in range(3)]
    print("K1 alone (host): %.3f ms per call" % ((time.perf_counter() - t0) * 100), flush=True)
    for name, fn in (("cascade", run_c), ("forward", run_f)):
        eager = timeit(fn, args.reps)
        ref = fn()["depth"].clone()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        try:
            with torch.cuda.graph(g):
                out = fn()
        except Exception as e:  # noqa
            print("%s: capture failed: %r" % (name, e), flush=True)
            continue
        g.replay(); torch.cuda.synchronize()
        same = torch.equal(out["depth"], ref)
        graph = timeit(g.replay, args.reps)
        print("%s: eager %.3f ms, graph replay %.3f ms, identical depth: %s" % (name, eager, graph, same), flush=True)
        del g, out
