"""GPU: vido_convert_raw (csrc/input_kernels.cu) -- Bayer RG -> BGR, u16 -> f32, u8 -> i32 -- bit-exact against the cv2 goldens at
the golden sizes and against the oracle's numpy restatement at the KAIST image size (1280 x 560) and the KITTI size."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import input_oracle as io_oracle  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(HERE, "golden", "input_golden.npz"))


def _convert(pkg, bayer=None, depth16=None, mask8=None, on_device=False):
    ref = bayer if bayer is not None else (depth16 if depth16 is not None else mask8)
    n, H, W = ref.shape
    ctx = pkg.Context(pkg.default_config(width=W, height=H, max_batch=1))
    d_bgr = torch.zeros((n, H, W, 3), dtype=torch.uint8, device="cuda")
    d_dep = torch.zeros((n, H, W), dtype=torch.float32, device="cuda")
    d_msk = torch.zeros((n, H, W), dtype=torch.int32, device="cuda")
    keep = []

    def src(a):
        if a is None:
            return None
        if not on_device:
            return a
        t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).cuda()
        keep.append(t)
        return t.data_ptr()
    ctx.convert_raw(n, d_bgr.data_ptr(), d_dep.data_ptr(), d_msk.data_ptr(), bayer=src(bayer), depth16=src(depth16), mask8=src(mask8))
    torch.cuda.synchronize()
    out = d_bgr.cpu().numpy(), d_dep.cpu().numpy(), d_msk.cpu().numpy()
    ctx.close()
    return out


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_demosaic_equals_cv2_golden(pkg, name):
    raw = GOLD["bayer_" + name]
    bgr, _, _ = _convert(pkg, bayer=raw[None])
    assert np.array_equal(bgr[0], GOLD["bgr_" + name])


@pytest.mark.parametrize("shape,on_device", [((560, 1280), False), ((375, 1242), True), ((67, 129), False)])
def test_full_size_conversions_equal_oracle(pkg, shape, on_device):
    rng = np.random.default_rng(7)
    n = 3
    raw = rng.integers(0, 256, (n,) + shape, dtype=np.uint8)
    d16 = rng.integers(0, 65536, (n,) + shape, dtype=np.uint16)
    m8 = rng.integers(0, 6, (n,) + shape, dtype=np.uint8)
    bgr, dep, msk = _convert(pkg, raw, d16, m8, on_device=on_device)
    for k in range(n):
        assert np.array_equal(bgr[k], io_oracle.bayer_rg2bgr(raw[k])), k
    assert np.array_equal(dep, io_oracle.depth_to_f32(d16))
    assert np.array_equal(msk, io_oracle.mask_to_i32(m8))


def _write_png(path, a):
    """minimal PNG writer (grey 8 / 16 bit, filter 0) so that the test does not depend on an image library"""
    import struct
    import zlib
    h, w = a.shape
    depth = 8 if a.dtype == np.uint8 else 16
    rows = a.astype(">u2").tobytes() if depth == 16 else a.tobytes()
    stride = w * depth // 8
    raw = b"".join(b"\x00" + rows[y * stride:(y + 1) * stride] for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def test_demo_binary_tracks_a_kaist_layout_sequence(tmp_path, pkg):
    """run_vido_slam (host/run_vido_slam.cc, the reference's demo/run_vido_slam.cc) on a directory in the KAIST layout: raw Bayer
    PNGs, .flo flow, 16-bit depth PNGs, 8-bit mask PNGs, vTimestampsImage.txt.  Its trajectory file equals what the Python
    binding returns for the same frames converted on the host with the oracle (Bayer -> BGR, u16 -> f32, u8 -> i32)."""
    import subprocess
    import synth
    root = os.path.dirname(HERE)
    cam, n = synth.SMALL, 6
    sc = synth.Scene(cam=cam, seed=31, flow_noise=0.1, depth_noise=0.01)
    seq = tmp_path / "seq"
    for d in ("image", "flow_image", "depth_image", "mask_image"):
        (seq / d).mkdir(parents=True)
    stamps = [1544590798702901234 + 100000000 * k for k in range(n)]
    with open(seq / "vTimestampsImage.txt", "w") as fh:
        fh.write("timestamp\n" + "".join(f"{s}\n" for s in stamps))
    names, times = pkg.load_kaist_timestamps(seq / "image")
    host = []
    for k in range(n):
        f = sc.frame(k)
        raw = f["gray"].numpy()
        d16 = np.clip(np.rint(f["depth_in"].numpy()), 0, 65535).astype(np.uint16)
        m8 = f["mask"].numpy().astype(np.uint8)
        flow = f["flow"].numpy().astype(np.float32)
        _write_png(seq / "image" / (names[k] + ".png"), raw)
        _write_png(seq / "depth_image" / (names[k] + ".png"), d16)
        _write_png(seq / "mask_image" / (names[k] + ".png"), m8)
        with open(seq / "flow_image" / (names[k] + ".flo"), "wb") as fh:
            fh.write(b"PIEH" + np.array([flow.shape[1], flow.shape[0]], np.int32).tobytes() + flow.tobytes())
        host.append(dict(image=io_oracle.bayer_rg2bgr(raw), depth=io_oracle.depth_to_f32(d16), flow=flow, mask=io_oracle.mask_to_i32(m8),
                         timestamp=float(times[k])))
    yaml = tmp_path / "cfg.yaml"
    yaml.write_text("%YAML:1.0\n" + "".join(f"Camera.{k}: {cam[k]}\n" for k in ("width", "height", "fx", "fy", "cx", "cy", "bf")) +
                    f"ChooseData: 2\nDepthMapFactor: 256.0\nimage_path: \"{seq / 'image'}\"\nimu_path: \"\"\nstart_index: 0\nslam_mode: 0\n")
    out = tmp_path / "res_"
    run = subprocess.run([os.path.join(root, "vido-slam_b200", "run_vido_slam"), str(yaml), str(out)], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-800:]
    assert "processing image idx --> 5" in run.stdout
    ini = np.loadtxt(str(out) + "initial_rgbd_new.txt")
    assert ini.shape == (n, 17)
    ctx = pkg.Context(pkg.default_config(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"],
                                         bf=cam["bf"], max_batch=1))
    for f in host:
        ctx.track_frames([f])
    P = ctx.map_poses().reshape(n, 16)
    assert np.abs(ini[:, 1:] - P).max() <= 1e-6 * max(np.abs(P).max(), 1.0)
    ctx.close()
