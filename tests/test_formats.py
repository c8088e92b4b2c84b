"""dmvsnet_b200/formats.py against files written and parsed by the live reference (tools/make_golden_formats.py): the writers
must produce the reference's bytes, the readers the reference's arrays.  CPU only."""
import os

import numpy as np
import pytest

from dmvsnet_b200 import formats as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "formats")


@pytest.fixture(scope="module")
def parsed():
    return np.load(os.path.join(GOLD, "parsed.npz"))


def _bytes(path):
    with open(path, "rb") as f:
        return f.read()


def test_save_pfm_writes_the_reference_bytes(parsed, tmp_path):
    F.save_pfm(str(tmp_path / "g.pfm"), parsed["gray_in"])
    F.save_pfm(str(tmp_path / "c.pfm"), parsed["color_in"], scale=2)
    assert _bytes(tmp_path / "g.pfm") == _bytes(os.path.join(GOLD, "gray.pfm"))
    assert _bytes(tmp_path / "c.pfm") == _bytes(os.path.join(GOLD, "color.pfm"))
    F.save_pfm(str(tmp_path / "g1.pfm"), parsed["gray_in"][:, :, None])          # H x W x 1 is greyscale too
    assert _bytes(tmp_path / "g1.pfm") == _bytes(os.path.join(GOLD, "gray.pfm"))
    with pytest.raises(Exception, match="float32"):
        F.save_pfm(str(tmp_path / "bad.pfm"), parsed["gray_in"].astype(np.float64))
    with pytest.raises(Exception, match="dimensions"):
        F.save_pfm(str(tmp_path / "bad.pfm"), np.zeros((2, 3, 2), np.float32))


def test_read_pfm_matches_the_reference(parsed, tmp_path):
    for name in ("gray", "color"):
        data, scale = F.read_pfm(os.path.join(GOLD, name + ".pfm"))
        assert np.array_equal(data, parsed[name]) and scale == float(parsed[name + "_scale"])
        assert np.array_equal(data, parsed[name + "_in"])                         # save -> read round trip, top row first
    big, scale = F.read_pfm(os.path.join(GOLD, "big_endian.pfm"))
    assert np.array_equal(big.astype(np.float32), parsed["big"]) and scale == 1.0
    (tmp_path / "x.pfm").write_bytes(b"P6\n1 1\n-1\n0000")
    with pytest.raises(Exception, match="Not a PFM"):
        F.read_pfm(str(tmp_path / "x.pfm"))
    (tmp_path / "y.pfm").write_bytes(b"Pf\n1x1\n-1\n0000")
    with pytest.raises(Exception, match="Malformed"):
        F.read_pfm(str(tmp_path / "y.pfm"))


def test_camera_and_pair_files(parsed, tmp_path):
    k, e = F.read_camera_parameters(os.path.join(GOLD, "00000000_cam.txt"))
    assert k.dtype == np.float32 and e.dtype == np.float32
    assert np.array_equal(k, parsed["intrinsics"]) and np.array_equal(e, parsed["extrinsics"])
    F.write_cam(str(tmp_path / "cam.txt"), parsed["cam_in"])
    assert _bytes(tmp_path / "cam.txt") == _bytes(os.path.join(GOLD, "00000000_cam.txt"))
    pairs = F.read_pair_file(os.path.join(GOLD, "pair.txt"))
    assert [p[0] for p in pairs] == parsed["pair_refs"].tolist()                  # view 1 has no source views: dropped
    assert [p[1] for p in pairs] == [[s for s in row if s >= 0] for row in parsed["pair_srcs"].tolist()]


def test_ply_round_trip_and_header(tmp_path):
    rng = np.random.RandomState(0)
    pts = rng.randn(11, 3)
    col = rng.randint(0, 256, (11, 3)).astype(np.uint8)
    path = str(tmp_path / "cloud.ply")
    F.write_ply(path, pts, col)
    raw = _bytes(path)
    head, body = raw.split(b"end_header\n")
    assert head.decode().splitlines() == ["ply", "format binary_little_endian 1.0", "element vertex 11", "property float x", "property float y",
                                          "property float z", "property uchar red", "property uchar green", "property uchar blue"]
    assert len(body) == 11 * 15
    p2, c2 = F.read_ply(path)
    assert np.array_equal(p2, pts.astype(np.float32)) and np.array_equal(c2, col)
    with pytest.raises(ValueError):
        F.write_ply(path, pts, col[:5])
    F.write_ply(path, np.zeros((0, 3)), np.zeros((0, 3), np.uint8))
    assert F.read_ply(path)[0].shape == (0, 3)
