import sys; sys.path.insert(0, "/root/repo")
import torch
from dmvsnet_b200 import MVSNet, synthetic as syn
from oracle import dmvs_oracle as O
for (b, n, h, w, natural) in ((2, 3, 64, 96, False), (1, 4, 96, 128, True), (2, 3, 64, 96, True)):
    nd, ratios = [16, 8, 8], [4, 2, 1]
    net = MVSNet(nd, ratios, inverse_depth=True)
    state = syn.randomise_regnet_state(net.state_dict(), seed=5)
    net.load_state_dict(state); net = net.cuda().eval()
    imgs = syn.make_images(h, w, n, b, seed=6, natural=natural)
    proj = syn.make_proj_matrices(h, w, n, b, num_stages=3)
    dv = syn.make_depth_values(b, 192, inverse=True)
    with torch.no_grad():
        got = net(imgs.cuda(), proj, dv.cuda())
        want = O.mvsnet_forward(imgs, proj, dv, state, nd, ratios, True)
    for k in ("depth", "photometric_confidence"):
        e = ((got[k].cpu() - want[k]).abs() / want[k].abs().clamp_min(1.0))
        print(b, n, h, w, natural, k, "max rel", float(e.max()), "p99.9", float(e.flatten().kthvalue(int(0.999 * e.numel()))[0]))
    for s in ("stage1", "stage2", "stage3"):
        e = ((got[s]["depth"].cpu() - want[s]["depth"]).abs() / want[s]["depth"].abs())
        print("   ", s, "depth max rel", float(e.max()))
