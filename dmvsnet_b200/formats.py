"""On-disk formats either side of the path: what ``Model.test`` writes after the forward and what the
depth-map fusion reads back and finally emits.

* PFM depth / confidence maps    - reference datasets/data_io.py:6-71 (``read_pfm`` / ``save_pfm``): same bytes on disk
* ``*_cam.txt`` camera files      - reference filter/pcd.py:52-63 (``read_camera_parameters``), tools.py:40-57 (``write_cam``)
* ``pair.txt`` view selection     - reference filter/pcd.py:66-78 (``read_pair_file``)
* PLY point clouds                - reference filter/pcd.py:349-361 (plyfile ``PlyData([PlyElement.describe(vertex_all, 'vertex')])``,
  binary little endian, x y z float + red green blue uchar); written with numpy alone, plyfile is not a dependency

Host-side byte shuffling only; nothing here touches the GPU.
"""
from __future__ import annotations

import re
import sys
from typing import List, Sequence, Tuple

import numpy as np

_PFM_DIMS = re.compile(rb"^(\d+)\s(\d+)\s$")


def read_pfm(filename: str) -> Tuple[np.ndarray, float]:
    """-> (data, scale): data [H,W] ('Pf') or [H,W,3] ('PF'), top row first (the file stores the bottom row first), in the
    file's byte order like the reference (``np.fromfile(f, '<f')``); scale is returned positive."""
    with open(filename, "rb") as f:
        magic = f.readline().rstrip()
        if magic not in (b"PF", b"Pf"):
            raise Exception("Not a PFM file.")
        dims = _PFM_DIMS.match(f.readline())
        if not dims:
            raise Exception("Malformed PFM header.")
        width, height = int(dims.group(1)), int(dims.group(2))
        scale = float(f.readline().rstrip())
        order = "<" if scale < 0 else ">"   # a negative scale marks little-endian samples
        data = np.fromfile(f, order + "f")
    shape = (height, width, 3) if magic == b"PF" else (height, width)
    return np.flipud(data.reshape(shape)), abs(scale)


def save_pfm(filename: str, image: np.ndarray, scale: float = 1) -> None:
    """float32 [H,W], [H,W,1] or [H,W,3]; rows are written bottom-up, the scale line carries the byte order's sign."""
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        magic = b"PF\n"
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        magic = b"Pf\n"
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    little = image.dtype.byteorder == "<" or (image.dtype.byteorder == "=" and sys.byteorder == "little")
    with open(filename, "wb") as f:
        f.write(magic)
        f.write(b"%d %d\n" % (image.shape[1], image.shape[0]))
        f.write(("%f\n" % (-scale if little else scale)).encode("utf-8"))
        np.flipud(image).tofile(f)


def read_camera_parameters(filename: str) -> Tuple[np.ndarray, np.ndarray]:
    """``*_cam.txt`` -> (intrinsics [3,3], extrinsics [4,4]) float32: lines 1-4 hold the extrinsic, lines 7-9 the intrinsic."""
    with open(filename) as f:
        lines = [ln.rstrip() for ln in f.readlines()]
    extrinsics = np.array(" ".join(lines[1:5]).split(), dtype=np.float32).reshape(4, 4)
    intrinsics = np.array(" ".join(lines[7:10]).split(), dtype=np.float32).reshape(3, 3)
    return intrinsics, extrinsics


def write_cam(filename: str, cam: np.ndarray) -> None:
    """cam [2,4,4]: cam[0] extrinsic, cam[1][:3,:3] intrinsic, cam[1][3] = (depth_min, interval, ndepth, depth_max) - the layout
    of the dataset's ``proj_matrices`` entries; text layout of tools.py:40-57 (every number followed by a blank)."""
    with open(filename, "w") as f:
        f.write("extrinsic\n")
        for row in cam[0]:
            f.write("".join(str(v) + " " for v in row) + "\n")
        f.write("\nintrinsic\n")
        for row in cam[1][:3]:
            f.write("".join(str(v) + " " for v in row[:3]) + "\n")
        f.write("\n" + " ".join(str(v) for v in cam[1][3]) + "\n")


def read_pair_file(filename: str) -> List[Tuple[int, List[int]]]:
    """``pair.txt`` -> [(ref_view, [src_view, ...]), ...]; reference views without source views are dropped."""
    out = []
    with open(filename) as f:
        n = int(f.readline())
        for _ in range(n):
            ref = int(f.readline().rstrip())
            src = [int(x) for x in f.readline().rstrip().split()[1::2]]   # "<count> id score id score ..."
            if src:
                out.append((ref, src))
    return out


_PLY_VERTEX = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])


def write_ply(filename: str, points: np.ndarray, colors: np.ndarray) -> None:
    """points [N,3] float, colors [N,3] uint8 -> the binary little-endian PLY plyfile writes for the reference's vertex element."""
    points = np.asarray(points)
    colors = np.asarray(colors)
    if points.ndim != 2 or points.shape[1] != 3 or colors.shape != points.shape:
        raise ValueError("points / colors must both be [N,3], got %s / %s" % (points.shape, colors.shape))
    v = np.empty(len(points), _PLY_VERTEX)
    v["x"], v["y"], v["z"] = points[:, 0], points[:, 1], points[:, 2]
    v["red"], v["green"], v["blue"] = colors[:, 0], colors[:, 1], colors[:, 2]
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % len(v))
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        v.tofile(f)


def read_ply(filename: str) -> Tuple[np.ndarray, np.ndarray]:
    """Inverse of ``write_ply`` (the vertex element only): -> (points [N,3] float32, colors [N,3] uint8)."""
    with open(filename, "rb") as f:
        n = None
        while True:
            line = f.readline()
            if not line:
                raise ValueError("PLY header not terminated")
            if line.startswith(b"element vertex"):
                n = int(line.split()[2])
            if line.strip() == b"end_header":
                break
        v = np.fromfile(f, _PLY_VERTEX, count=n)
    return np.stack([v["x"], v["y"], v["z"]], 1), np.stack([v["red"], v["green"], v["blue"]], 1)
