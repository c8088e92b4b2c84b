"""Ad-hoc: where does cascade_lin deviate? (run on the GPU box)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from cases import CASES, case_inputs, case_state
from oracle import dmvs_oracle as O
from dmvsnet_b200 import ops, MVSNet

name = sys.argv[1] if len(sys.argv) > 1 else "cascade_lin"
case = CASES[name]
z = np.load(os.path.join(ROOT, "tests/golden/%s.npz" % name)); gold = {k: torch.from_numpy(z[k]) for k in z.files}
inp = case_inputs(case)
feats = inp["features"]
print("feature checksums", [float(f["stage1"].double().sum()) for f in feats])
hyp = gold["s1_depth_values"]
rt_here = ops.relative_projections(inp["proj"]["stage1"])
print("rt here vs fixture max abs diff", float((rt_here - gold["s1_rt"]).abs().max()), "rel", float(((rt_here - gold["s1_rt"]).abs() / gold["s1_rt"].abs().clamp_min(1e-6)).max()))
cost_oracle_here = O.warp_corr([f["stage1"] for f in feats], inp["proj"]["stage1"], hyp)
def rel(a, b): return float((a - b).abs().max() / b.abs().max())
print("oracle(here) vs golden cost:", rel(cost_oracle_here, gold["s1_cost"]))
if torch.cuda.is_available():
    g = ops.warp_corr([f["stage1"].cuda() for f in feats], gold["s1_rt"].cuda(), hyp.cuda()).cpu()
    print("gpu(fixture rt) vs golden:", rel(g, gold["s1_cost"]), " vs oracle(here):", rel(g, cost_oracle_here))
    g2 = ops.warp_corr([f["stage1"].cuda() for f in feats], rt_here.cuda(), hyp.cuda()).cpu()
    print("gpu(rt here) vs golden:", rel(g2, gold["s1_cost"]), " vs oracle(here):", rel(g2, cost_oracle_here))
    d = (g - gold["s1_cost"]).abs()
    idx = torch.nonzero(d > 0.5 * d.max())[:10]
    print("worst locations (b,g,d,y,x):", idx.tolist())
