"""Test infrastructure only (oracle): numpy restatement of the demo's input staging, vido_slam/demo/run_vido_slam.cc:114-122.
OpenCV is a third-party dependency that is not part of /root/reference (linked 3.4); the functions below
follow its published algorithms and are pinned against cv2 4.13 run in the authoring container (tests/golden/make_input_golden.py
-> tests/golden/input_golden.npz + the small files next to it)."""
import struct
import zlib

import numpy as np


def bayer_rg2bgr(raw):
    """cv::cvtColor(raw, COLOR_BayerRG2BGR), 8 bit: bilinear demosaic (modules/imgproc/src/demosaicing.cpp, Bayer2RGB_); the first /
    last column copy their inner neighbour, then the first / last row copy theirs."""
    H, W = raw.shape
    P = np.pad(raw.astype(np.int32), 1, mode="edge")
    c = P[1:-1, 1:-1]
    up, dn, lf, rt = P[:-2, 1:-1], P[2:, 1:-1], P[1:-1, :-2], P[1:-1, 2:]
    cross = (up + dn + lf + rt + 2) >> 2
    diag = (P[:-2, :-2] + P[:-2, 2:] + P[2:, :-2] + P[2:, 2:] + 2) >> 2
    hor, ver = (lf + rt + 1) >> 1, (up + dn + 1) >> 1
    yy, xx = np.mgrid[0:H, 0:W]
    out = np.zeros((H, W, 3), np.int32)
    for (py, px), f in (((0, 0), (c, cross, diag)), ((0, 1), (hor, c, ver)), ((1, 0), (ver, c, hor)), ((1, 1), (diag, cross, c))):
        sel = (yy % 2 == py) & (xx % 2 == px)
        for ch in range(3):
            out[..., ch][sel] = f[ch][sel]
    out[:, 0] = out[:, 1]; out[:, -1] = out[:, -2]
    out[0] = out[1]; out[-1] = out[-2]
    return out.astype(np.uint8)


def depth_to_f32(d16):
    return d16.astype(np.float32)      # Mat::convertTo(CV_32F)


def mask_to_i32(m8):
    return m8.astype(np.int32)         # Mat::convertTo(CV_32SC1)


def read_png(path):
    """cv::imread(IMREAD_UNCHANGED) for non-interlaced grey / RGB(A) PNG, channels in file order (PNG specification, zlib)"""
    f = open(path, "rb").read()
    assert f[:8] == b"\x89PNG\r\n\x1a\n"
    pos, z = 8, b""
    while pos < len(f):
        n, typ = struct.unpack(">I4s", f[pos:pos + 8])
        d = f[pos + 8:pos + 8 + n]
        if typ == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", d)
            assert interlace == 0
        elif typ == b"IDAT":
            z += d
        pos += 12 + n
    ch = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    bpp = ch * depth // 8
    row = bpp * w
    raw = np.frombuffer(zlib.decompress(z), np.uint8).reshape(h, row + 1)
    out = np.zeros((h, row), np.uint8)
    for y in range(h):
        ft, src = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        up = out[y - 1].astype(np.int32) if y else np.zeros(row, np.int32)
        cur = np.zeros(row, np.int32)
        if ft == 0:
            cur = src
        elif ft == 2:
            cur = (src + up) & 255
        else:
            for i in range(row):
                a = cur[i - bpp] if i >= bpp else 0
                b = up[i]
                cc = up[i - bpp] if i >= bpp else 0
                if ft == 1:
                    pr = a
                elif ft == 3:
                    pr = (a + b) >> 1
                else:
                    p = a + b - cc
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - cc)
                    pr = a if (pa <= pb and pa <= pc) else (b if pb <= pc else cc)
                cur[i] = (src[i] + pr) & 255
        out[y] = cur
    if depth == 16:
        out = out.reshape(h, w * ch, 2)
        a = (out[..., 0].astype(np.uint16) << 8) | out[..., 1]
    else:
        a = out
    a = a.reshape(h, w, ch)
    return a[:, :, 0] if ch == 1 else a


def read_flo(path):
    """cv::optflow::readOpticalFlow (Middlebury .flo)"""
    f = open(path, "rb").read()
    assert f[:4] == b"PIEH"
    w, h = struct.unpack("<ii", f[4:12])
    return np.frombuffer(f[12:12 + 8 * w * h], np.float32).reshape(h, w, 2).copy()
