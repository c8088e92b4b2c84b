"""A tiny scan folder + what the LIVE reference's ``filter_depth`` makes of it (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_scene.py     # -> tests/golden/fusion_scene/ , fusion_scene_expected.npz

The scene (4 views, 32x48: cams, jpg images, PFM depth / confidence maps, pair.txt) is written with dmvsnet_b200.formats - whose
writers are byte-identical to the reference's (tests/test_formats.py).  Then the reference's own ``filter_depth`` runs on a copy
of it, twice: filter/pcd.py (fixed thresholds) and filter/dypcd_tanks.py (dynamic thresholds).  plyfile is absent here: its two
entry points are replaced by a recorder that keeps the vertex array the reference hands to it; ``.cuda()`` is neutralised (no GPU)
so pcd.py's torch code runs on the CPU.  Recorded: vertices, colours and the three masks per reference view.
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_fusion import fusion_case  # noqa: E402
from dmvsnet_b200 import formats  # noqa: E402

SCENE = os.path.join(ROOT, "tests", "golden", "fusion_scene")
ARGS = dict(ndepths=[48, 32, 8], conf=[0.1, 0.15, 0.2], thres_view=2, dist_base=1 / 4, rel_diff_base=1 / 1300, display=False)


def write_scene(folder, h=32, w=48, views=4):
    from PIL import Image
    depths, ks, es = fusion_case(h, w, views, seed=4)
    rng = np.random.RandomState(7)
    for sub in ("cams", "images", "depth_est", "confidence"):
        os.makedirs(os.path.join(folder, sub), exist_ok=True)
    with open(os.path.join(folder, "pair.txt"), "w") as f:
        f.write("%d\n" % views)
        for v in range(views):
            others = [u for u in range(views) if u != v]
            f.write("%d\n%d %s\n" % (v, len(others), " ".join("%d %.2f" % (u, 100.0 - u) for u in others)))
    for v in range(views):
        cam = np.zeros((2, 4, 4), np.float32)
        cam[0] = es[v].numpy()
        cam[1, :3, :3] = ks[v].numpy()
        cam[1, 3] = np.array([425.0, 2.5, 192, 902.5], np.float32)
        formats.write_cam(os.path.join(folder, "cams/{:0>8}_cam.txt".format(v)), cam)
        Image.fromarray(rng.randint(0, 256, (h, w, 3)).astype(np.uint8)).save(os.path.join(folder, "images/{:0>8}.jpg".format(v)), quality=95)
        formats.save_pfm(os.path.join(folder, "depth_est/{:0>8}.pfm".format(v)), depths[v].numpy())
        formats.save_pfm(os.path.join(folder, "confidence/{:0>8}.pfm".format(v)), rng.rand(h, w).astype(np.float32))
    for stage in (1, 2):   # per-stage confidences for view 0 only: both branches of pcd.py:267-271
        formats.save_pfm(os.path.join(folder, "confidence/{:0>8}_stage{}.pfm".format(0, stage)), rng.rand(h, w).astype(np.float32))


def main():
    from PIL import Image
    captured = {}

    class _Any:
        def __init__(self, *a, **k): pass
        def __getattr__(self, n): return _Any()
        def __call__(self, *a, **k): return _Any()

    class PlyElement:
        @staticmethod
        def describe(arr, name): return arr

    class PlyData:
        def __init__(self, els): self.els = els
        def write(self, fn): captured[os.path.basename(fn)] = self.els[0]

    m = types.ModuleType("plyfile"); m.PlyData, m.PlyElement = PlyData, PlyElement; sys.modules["plyfile"] = m
    for name in ("tomlkit", "yacs", "yacs.config"):
        m = types.ModuleType(name); m.value = None; m.CfgNode = _Any; sys.modules[name] = m
    m = types.ModuleType("filter.tank_test_config"); m.tank_cfg = _Any(); sys.modules["filter.tank_test_config"] = m
    sys.path.insert(0, "/root/reference")
    import filter.pcd as pcd
    import filter.dypcd_tanks as dy
    torch.Tensor.cuda = lambda self, *a, **k: self
    dy.np.bool = np.bool_  # dypcd's save_mask asserts on the alias

    if os.path.exists(SCENE):
        shutil.rmtree(SCENE)
    write_scene(SCENE)
    args = types.SimpleNamespace(**ARGS)
    out = {}
    for tag, mod in (("static", pcd), ("dynamic", dy)):
        with tempfile.TemporaryDirectory() as tmp:
            work = os.path.join(tmp, "scan")
            shutil.copytree(SCENE, work)
            mod.filter_depth(args, work, work, work, os.path.join(tmp, tag + ".ply"))
            v = captured[tag + ".ply"]
            out[tag + "_xyz"] = np.stack([v["x"], v["y"], v["z"]], 1)
            out[tag + "_rgb"] = np.stack([v["red"], v["green"], v["blue"]], 1)
            for view in range(4):
                for kind in ("photo", "geo", "final"):
                    out["%s_mask_%s_%d" % (tag, kind, view)] = np.array(Image.open(os.path.join(work, "mask/{:0>8}_{}.png".format(view, kind)))) > 0
        print(tag, "points:", len(v), "final-mask fractions:", [float(out["%s_mask_final_%d" % (tag, k)].mean()) for k in range(4)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fusion_scene_expected.npz"), **out)


if __name__ == "__main__":
    main()
