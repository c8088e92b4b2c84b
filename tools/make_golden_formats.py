"""Golden files for dmvsnet_b200/formats.py from the LIVE reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_formats.py        # -> tests/golden/formats/

The reference's own writers produce the files (datasets/data_io.py save_pfm, tools.py write_cam), the reference's own readers
(data_io.read_pfm, filter/pcd.py read_camera_parameters / read_pair_file) produce the parsed arrays next to them.  filter/pcd.py is
imported with its unrelated absent dependencies stubbed, like tools/make_golden_fusion.py does.  plyfile is not installed here, so
the PLY writer has no reference-made golden file (its header layout is plyfile's documented output for this vertex dtype).
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "formats")


def main():
    class _Any:
        def __init__(self, *a, **k): pass
        def __getattr__(self, n): return _Any()
        def __call__(self, *a, **k): return _Any()
    for name in ("plyfile", "tomlkit", "yacs", "yacs.config"):
        m = types.ModuleType(name); m.PlyData = m.PlyElement = m.value = None; m.CfgNode = _Any; sys.modules[name] = m
    m = types.ModuleType("filter.tank_test_config"); m.tank_cfg = _Any(); sys.modules["filter.tank_test_config"] = m
    sys.path.insert(0, "/root/reference")
    from datasets import data_io
    import filter.pcd as pcd
    import tools as ref_tools

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.RandomState(3)
    gray = (400 + 500 * rng.rand(5, 7)).astype(np.float32)
    color = rng.rand(4, 3, 3).astype(np.float32)
    data_io.save_pfm(os.path.join(OUT, "gray.pfm"), gray)
    data_io.save_pfm(os.path.join(OUT, "color.pfm"), color, scale=2)
    with open(os.path.join(OUT, "big_endian.pfm"), "wb") as f:        # a big-endian file as other tools write it
        f.write(b"Pf\n7 5\n1.000000\n")
        np.flipud(gray).astype(">f4").tofile(f)
    cam = np.zeros((2, 4, 4), np.float32)
    cam[0] = np.eye(4, dtype=np.float32)
    cam[0, :3, :] = rng.randn(3, 4).astype(np.float32)
    cam[1, :3, :3] = np.array([[2892.33, 0, 823.205], [0, 2883.175, 619.071], [0, 0, 1]], np.float32)
    cam[1, 3] = np.array([425.0, 2.5, 192, 902.5], np.float32)
    ref_tools.write_cam(os.path.join(OUT, "00000000_cam.txt"), cam)
    with open(os.path.join(OUT, "pair.txt"), "w") as f:
        f.write("3\n0\n3 10 2346.41 1 2036.53 9 1243.89\n1\n0\n2\n2 0 1.5 1 0.25\n")
    k, e = pcd.read_camera_parameters(os.path.join(OUT, "00000000_cam.txt"))
    pairs = pcd.read_pair_file(os.path.join(OUT, "pair.txt"))
    g, gs = data_io.read_pfm(os.path.join(OUT, "gray.pfm"))
    c, cs = data_io.read_pfm(os.path.join(OUT, "color.pfm"))
    b, bs = data_io.read_pfm(os.path.join(OUT, "big_endian.pfm"))
    np.savez(os.path.join(OUT, "parsed.npz"), gray_in=gray, color_in=color, cam_in=cam, gray=np.ascontiguousarray(g), gray_scale=gs,
             color=np.ascontiguousarray(c), color_scale=cs, big=np.ascontiguousarray(b).astype(np.float32), big_scale=bs, intrinsics=k,
             extrinsics=e, pair_refs=np.array([p[0] for p in pairs]), pair_srcs=np.array([p[1] + [-1] * (3 - len(p[1])) for p in pairs]))
    print("wrote", sorted(os.listdir(OUT)), "pairs:", pairs)


if __name__ == "__main__":
    main()
