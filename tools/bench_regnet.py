"""Ad-hoc (GPU box): the regularisation nets on the DTU stage grids, one call per (stage, net), CUDA-event time per call with the
depth-tap-folded kernels (csrc/conv_kf.cu) on and off; the logits of both settings are compared.

    python tools/bench_regnet.py [--reps 10]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dmvsnet_b200 import MVSNet, _native, ops, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--knob", default="kf")
ap.add_argument("--values", default="0,1")
ap.add_argument("--set", default="", help="other knobs, e.g. regnet_streams=0,tc2_pdl=0")
ap.add_argument("--stages", default="1,2,3")
args = ap.parse_args()
dev = torch.device("cuda:0")
lib = _native.load()
def apply_set():
    for kv in filter(None, args.set.split(",")):
        k_, v_ = kv.split("=")
        assert lib.dmvs_debug_set(k_.encode(), int(v_)) == 0
apply_set()
VALUES = [int(v) for v in args.values.split(",")]
STAGES = [int(v) for v in args.stages.split(",")]
net = MVSNet([48, 32, 8], [4, 2, 1], inverse_depth=True)
net.load_state_dict(syn.randomise_regnet_state(net.state_dict(), seed=0))
net = net.to(dev).eval()


def cells_of(vol):
    b, _, d, h, w = vol.shape
    hi = vol.to(torch.float16)
    lo = (vol - hi.float()).to(torch.float16)
    vox = torch.stack((hi[:, 0], hi[:, 1], lo[:, 0], lo[:, 1]), dim=-1)
    cells = torch.zeros(b, d, h, w + 1, 8, dtype=torch.float16, device=vol.device)
    cells[:, :, :, 1:, 0:4] = vox
    cells[:, :, :, :w, 4:8] = vox
    return cells.view(torch.int32).reshape(b, d, h, w + 1, 4)


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


total = {v: 0.0 for v in VALUES}
with torch.no_grad():
    for stage, (d, h, w) in enumerate([(48, 296, 400), (32, 592, 800), (8, 1184, 1600)]):
        if stage + 1 not in STAGES:
            continue
        for refine in (False, True):
            dd = 4 if refine else d
            mod = (net.cost_regularization_refine if refine else net.cost_regularization)[stage]
            vol = torch.randn(1, 2, dd, h, w, device=dev, generator=torch.Generator(device=dev).manual_seed(stage))
            cells = cells_of(vol)
            del vol
            pack = mod.packed()
            out = {}
            for v in VALUES:
                print('  running stage %d refine=%d %s=%d' % (stage + 1, refine, args.knob, v), flush=True)
                assert lib.dmvs_debug_set(args.knob.encode(), v) == 0
                apply_set()
                logits = torch.empty(1, 4, dd, h, w, device=dev)
                ms = timeit(lambda: ops.regnet_forward(pack, None, cost_cells=cells, out=logits), args.reps)
                out[v] = logits
                total[v] += ms
                print("stage %d %-6s D=%2d %4dx%4d  %s=%d  %.3f ms" % (stage + 1, "refine" if refine else "main", dd, h, w, args.knob, v, ms), flush=True)
            if len(VALUES) > 1:
                diff = (out[VALUES[0]] - out[VALUES[1]]).abs().max().item() / out[VALUES[0]].abs().max().item()
                print("    max |logit difference| / max |logit| = %.2e" % diff, flush=True)
            del out, logits, cells
            torch.cuda.empty_cache()
    lib.dmvs_debug_set(args.knob.encode(), 1)
print("sum over the calls: " + ", ".join("%s=%d %.3f ms" % (args.knob, v, total[v]) for v in VALUES))
