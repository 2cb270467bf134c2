"""CPU: the demo's on-disk formats (vido_slam/demo/run_vido_slam.cc:14-66,114-122) -- host/InputDecode.cc against files written
by cv2's encoders (tests/golden/make_input_golden.py), against the oracle's independent zlib restatement, and the Bayer oracle
against what cv2.cvtColor(COLOR_BayerRG2BGR) produced."""
import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import input_oracle as io_oracle  # noqa: E402

pkg = importlib.import_module("vido-slam_b200")
GOLD = np.load(os.path.join(HERE, "golden", "input_golden.npz"))
G = lambda name: os.path.join(HERE, "golden", name)


@pytest.mark.parametrize("fname,key", [("input_g8.png", "g8"), ("input_g16.png", "g16"), ("input_g16_smooth.png", "g16s")])
def test_grey_png_like_imread_unchanged(fname, key):
    a = pkg.read_png(G(fname))
    assert a.dtype == GOLD[key].dtype and np.array_equal(a, GOLD[key])
    assert np.array_equal(io_oracle.read_png(G(fname)), GOLD[key])


def test_colour_png_channels_in_file_order():
    a = pkg.read_png(G("input_c8.png"))
    assert np.array_equal(a[:, :, ::-1], GOLD["c8_bgr"])       # cv2 wrote BGR; the file holds RGB
    assert np.array_equal(io_oracle.read_png(G("input_c8.png")), a)


def test_flo_like_read_optical_flow():
    f = pkg.read_flo(G("input_flow.flo"))
    assert f.dtype == np.float32 and np.array_equal(f, GOLD["flow"])
    assert np.array_equal(io_oracle.read_flo(G("input_flow.flo")), GOLD["flow"])


def test_bad_files_are_rejected(tmp_path):
    p = tmp_path / "x.png"
    p.write_bytes(b"not a png at all")
    with pytest.raises(pkg.VidoError):
        pkg.read_png(p)
    q = tmp_path / "t.png"
    q.write_bytes(open(G("input_g16.png"), "rb").read()[:200])   # truncated
    with pytest.raises(pkg.VidoError):
        pkg.read_png(q)
    with pytest.raises(pkg.VidoError):
        pkg.read_flo(G("input_g8.png"))
    with pytest.raises(pkg.VidoError):
        pkg.read_png(tmp_path / "missing.png")


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_bayer_oracle_equals_cv2(name):
    assert np.array_equal(io_oracle.bayer_rg2bgr(GOLD["bayer_" + name]), GOLD["bgr_" + name])


def test_kaist_loaders_follow_the_demo(tmp_path):
    """LoadIMU: '#' lines skipped, t = col 0 / 1e9, gyro = cols 8-10, acc = cols 11-13 (:14-45); LoadKaistImg: header skipped,
    stem = the first 19 characters of the printed number (:47-66)"""
    img = tmp_path / "seq" / "image"
    img.mkdir(parents=True)
    rows = []
    for k in range(5):
        t = 1544590798702901234 + 5000000 * k
        rows.append([t] + [0.0] * 7 + [0.1 * k, 0.2 * k, 0.3 * k, 9.0 + k, -1.0 * k, 0.5 * k] + [0.0] * 3)
    with open(tmp_path / "imu.csv", "w") as fh:
        fh.write("# timestamp, q, ..., gyro, acc\n")
        for r in rows:
            fh.write(",".join(repr(v) if isinstance(v, float) else str(v) for v in r) + "\n")
    a = pkg.load_kaist_imu(tmp_path / "imu.csv")
    assert a.shape == (5, 7)
    assert np.allclose(a[:, 0], [r[0] / 1e9 for r in rows], rtol=0, atol=1e-6)
    assert np.allclose(a[:, 1:4], [[np.float32(r[11]), np.float32(r[12]), np.float32(r[13])] for r in rows])
    assert np.allclose(a[:, 4:7], [[np.float32(r[8]), np.float32(r[9]), np.float32(r[10])] for r in rows])
    with open(tmp_path / "seq" / "vTimestampsImage.txt", "w") as fh:
        fh.write("timestamps\n1544590798702901234\n1544590798802901234\n")
    names, t = pkg.load_kaist_timestamps(img)
    assert len(names) == 2 and all(len(n) == 19 and n.isdigit() for n in names)
    assert names[0][:15] == "154459079870290"          # long double -> to_string keeps ~18-19 significant digits
    assert np.allclose(t, [1544590798.702901234, 1544590798.802901234], rtol=0, atol=1e-6)
