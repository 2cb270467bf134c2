"""CPU: the per-thread bodies of the descriptor kernels (vido-slam_b200/csrc/desc_device.h -- the source the __global__ wrappers
call) executed thread by thread on the host over the product's launch grids and the product's pitched pyramid layout
(tests/desc_emul.cc), bit for bit against the oracle and the cv2 goldens.  The GPU tests (test_zz_desc_gpu.py) repeat the same
comparisons through the C-ABI on the device."""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "desc_golden.npz")


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(tempfile.mkdtemp(prefix="vido_emul_"), "libdesc_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas", "-o", out,
                           os.path.join(ROOT, "tests", "desc_emul.cc")])
    lib = C.CDLL(out)
    lib.emul_hamming_part_bytes.restype = C.c_longlong
    return lib


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pattern():
    txt = open(os.path.join(ROOT, "include", "vido_orb_pattern.h")).read()
    body = txt[txt.index("{") + 1:txt.rindex("}")]
    v = np.array([int(t) for t in body.replace("\n", " ").split(",") if t.strip()], np.int8)
    assert v.size == 1024
    return v


class Layout:
    """orb_setup's pyramid layout (orb_kernels.cu): pitch = width rounded up to 64, frame stride rounded up to 256, levels back to back"""

    def __init__(self, sizes, nframes, scale):
        self.n = len(sizes)
        self.w = np.array([s[0] for s in sizes], np.int32)
        self.h = np.array([s[1] for s in sizes], np.int32)
        self.pitch = ((self.w + 63) // 64 * 64).astype(np.int32)
        self.fs = ((self.pitch.astype(np.int64) * self.h + 255) // 256 * 256).astype(np.int64)
        self.base = np.zeros(self.n, np.int64)
        off = 0
        for l in range(self.n):
            self.base[l] = off
            off += self.fs[l] * nframes
        self.bytes = int(off)
        self.nframes = nframes
        self.scale = np.asarray(scale, np.float32)

    def pack(self, pyramids, fill):
        if fill is None:   # random padding: bytes outside the level rectangles must not matter
            buf = np.random.default_rng(99).integers(0, 256, self.bytes, dtype=np.uint8)
        else:
            buf = np.full(self.bytes, fill, np.uint8)
        for f, pyr in enumerate(pyramids):
            for l, lvl in enumerate(pyr):
                v = buf[self.base[l] + f * self.fs[l]: self.base[l] + f * self.fs[l] + self.pitch[l] * self.h[l]].reshape(self.h[l], self.pitch[l])
                v[:, :self.w[l]] = lvl
        return buf

    def level(self, buf, f, l):
        v = buf[self.base[l] + f * self.fs[l]: self.base[l] + f * self.fs[l] + self.pitch[l] * self.h[l]].reshape(self.h[l], self.pitch[l])
        return v[:, :self.w[l]]


def run_blur(emul, lay, buf, version=1):
    out = np.full(lay.bytes, 0xEE, np.uint8)
    emul.emul_blur_v(_p(buf), _p(out), lay.n, _p(lay.w), _p(lay.h), _p(lay.pitch), _p(lay.base), _p(lay.fs), lay.nframes, version)
    return out


@pytest.mark.parametrize("version", [1, 2])
def test_blur_threads_match_cv2_and_oracle(emul, gold, version):
    p = ol.default_orb_params()
    imgs = [gold["kitti_scene_img"], gold["kitti_scene_next_img"]]
    w, h, s = ol.level_sizes(1242, 375, p)
    lay = Layout(list(zip(w, h)), 2, s)
    pyrs = [ol.orb_pyramid(im, p) for im in imgs]
    buf = lay.pack(pyrs, None)
    out = run_blur(emul, lay, buf, version)
    for f, name in enumerate(["kitti_scene", "kitti_scene_next"]):
        for l in range(lay.n):
            got = np.ascontiguousarray(lay.level(out, f, l))
            assert hashlib.sha256(got.tobytes()).hexdigest() == str(gold[f"{name}_blur_sha"][l]), (f, l)   # = cv2.GaussianBlur
    # nothing outside the level rectangles was written (row padding, frame padding)
    mask = np.ones(lay.bytes, bool)
    for f in range(2):
        for l in range(lay.n):
            v = mask[lay.base[l] + f * lay.fs[l]: lay.base[l] + f * lay.fs[l] + lay.pitch[l] * lay.h[l]].reshape(lay.h[l], lay.pitch[l])
            v[:, :lay.w[l]] = False
    assert (out[mask] == 0xEE).all()


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("size", [(8, 8), (9, 11), (13, 8), (64, 17), (67, 35), (131, 97), (333, 211), (12, 64)])
def test_blur_threads_odd_sizes(emul, size, version):
    """widths that are not multiples of 4, levels narrower than the vector path, heights that are not multiples of the strip"""
    rng = np.random.default_rng(size[0] * 1000 + size[1])
    w, h = size
    imgs = [rng.integers(0, 256, (h, w), dtype=np.uint8) for _ in range(3)]
    lay = Layout([(w, h)], 3, [1.0])
    out = run_blur(emul, lay, lay.pack([[im] for im in imgs], None), version)
    for f in range(3):
        assert np.array_equal(lay.level(out, f, 0), ol.gauss7(imgs[f])), f


def test_rbrief_threads_match_oracle_and_cv2(emul, gold):
    p = ol.default_orb_params()
    names = ["kitti_scene", "kitti_scene_next"]
    imgs = [gold[f"{n}_img"] for n in names]
    w, h, s = ol.level_sizes(1242, 375, p)
    lay = Layout(list(zip(w, h)), 2, s)
    blurred = run_blur(emul, lay, lay.pack([ol.orb_pyramid(im, p) for im in imgs], 0))
    cap = 2564
    kps = np.zeros((2, cap), ol.KP_DTYPE)
    nkp = np.zeros(2, np.int32)
    want = []
    for f, im in enumerate(imgs):
        k, d = ol.orb_extract_describe(im, p)
        kps[f, :len(k)] = k
        nkp[f] = len(k)
        want.append(d)
    desc = np.full((2, cap, 32), 0xEE, np.uint8)
    pat = pattern()
    emul.emul_rbrief(_p(blurred), lay.n, _p(lay.w), _p(lay.h), _p(lay.pitch), _p(lay.base), _p(lay.fs), _p(lay.scale), _p(kps), _p(nkp),
                     2, cap, _p(pat), _p(desc))
    for f in range(2):
        assert np.array_equal(desc[f, :nkp[f]], want[f]), f          # the oracle describes at level coordinates before scaling
        assert (desc[f, nkp[f]:] == 0xEE).all()                      # rows beyond the count stay untouched
        bits = np.unpackbits(desc[f, :nkp[f]] ^ gold[f"{names[f]}_desc_cv"], axis=1).sum(1)
        assert bits.mean() < 1.5                                     # cv2.ORB's own smoother differs by a grey level here and there


def test_rbrief_threads_foreign_key_points(emul):
    """key points that do not come from the extraction: next to the border (reads outside the level return 0, horizontal overshoot
    lands in the neighbouring row), every octave, arbitrary angles"""
    rng = np.random.default_rng(5)
    sizes = [(160, 120), (133, 100), (111, 83)]
    scale = np.array([1.0, 1.2, 1.44], np.float32)
    lay = Layout(sizes, 1, scale)
    lv = [rng.integers(0, 256, (h, w), dtype=np.uint8) for w, h in sizes]
    buf = lay.pack([lv], None)
    n = 600
    kps = np.zeros((1, n), ol.KP_DTYPE)
    octave = rng.integers(0, 3, n)
    lx = np.array([rng.integers(0, sizes[o][0]) for o in octave]).astype(np.float32)
    ly = np.array([rng.integers(0, sizes[o][1]) for o in octave]).astype(np.float32)
    kps["octave"][0] = octave
    kps["x"][0] = np.where(octave > 0, lx * scale[octave], lx)       # the extraction's scaling (float32 product)
    kps["y"][0] = np.where(octave > 0, ly * scale[octave], ly)
    kps["angle"][0] = rng.uniform(0, 360, n).astype(np.float32)
    nkp = np.array([n], np.int32)
    desc = np.zeros((1, n, 32), np.uint8)
    pat = pattern()
    emul.emul_rbrief(_p(buf), lay.n, _p(lay.w), _p(lay.h), _p(lay.pitch), _p(lay.base), _p(lay.fs), _p(lay.scale), _p(kps), _p(nkp), 1, n,
                     _p(pat), _p(desc))
    for o in range(3):
        sel = octave == o
        want = ol.describe_level(lv[o], lx[sel], ly[sel], kps["angle"][0][sel])
        assert np.array_equal(desc[0][sel], want), o


def run_hamming(emul, q, nq, t, nt, qcap):
    npairs = len(nq)
    part = np.zeros(emul.emul_hamming_part_bytes(npairs, qcap) // 4, np.int32)
    bi = np.full((npairs, qcap), -7, np.int32); bd = np.full((npairs, qcap), -7, np.int32); sd = np.full((npairs, qcap), -7, np.int32)
    emul.emul_hamming(_p(q), C.c_longlong(q.strides[0]), _p(nq), _p(t), C.c_longlong(t.strides[0]), _p(nt), npairs, qcap, _p(part),
                      _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


def test_hamming_threads_match_cv2_and_oracle(emul, gold):
    q0, t0 = gold["kitti_scene_desc_cv"], gold["kitti_scene_next_desc_cv"]
    qcap = 2600
    q = np.zeros((2, qcap, 32), np.uint8); t = np.zeros((2, 2700, 32), np.uint8)
    q[0, :len(q0)] = q0; t[0, :len(t0)] = t0
    q[1, :len(t0)] = t0; t[1, :len(q0)] = q0                         # second pair: the other direction
    nq = np.array([len(q0), len(t0)], np.int32); nt = np.array([len(t0), len(q0)], np.int32)
    bi, bd, sd = run_hamming(emul, q, nq, t, nt, qcap)
    assert np.array_equal(bd[0, :len(q0)], gold["match_best_dist"])
    assert np.array_equal(sd[0, :len(q0)], gold["match_second_dist"])
    assert np.array_equal(bi[0, :len(q0)], gold["match1_best_idx"])
    obi, obd, osd = ol.hamming_match(t0, q0)
    assert np.array_equal(bi[1, :len(t0)], obi) and np.array_equal(bd[1, :len(t0)], obd) and np.array_equal(sd[1, :len(t0)], osd)
    assert (bi[0, len(q0):] == -7).all() and (bi[1, len(t0):] == -7).all()


@pytest.mark.parametrize("nt", [0, 1, 2, 7, 8, 9, 63, 200])
def test_hamming_threads_small_and_tied_sets(emul, nt):
    """train sets smaller than the number of chunks, empty chunks, duplicated descriptors (ties: lowest index, second = best)"""
    rng = np.random.default_rng(nt)
    q = rng.integers(0, 256, (1, 40, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (1, max(nt, 1), 32), dtype=np.uint8)
    if nt >= 8:
        t[0, nt - 1] = t[0, 0]                                       # a duplicate in the last chunk
        q[0, 3] = t[0, 0]
    bi, bd, sd = run_hamming(emul, q, np.array([33], np.int32), t, np.array([nt], np.int32), 40)
    obi, obd, osd = ol.hamming_match(q[0, :33], t[0, :nt])
    assert np.array_equal(bi[0, :33], obi) and np.array_equal(bd[0, :33], obd) and np.array_equal(sd[0, :33], osd)
    if nt >= 8:
        assert bi[0, 3] == 0 and bd[0, 3] == 0 and sd[0, 3] == 0


def test_kernel_bodies_under_address_sanitizer(tmp_path):
    """the same per-thread bodies walked over their launch grids on heap buffers of exactly the product's sizes, under
    AddressSanitizer and with the 4-byte alignment of every vector access asserted (tests/desc_emul_asan.cc): nine image sizes,
    key points on and outside the border, out-of-range octaves, the matcher's partial / output arrays without slack"""
    exe = str(tmp_path / "desc_asan")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address", "-fno-omit-frame-pointer", "-ffp-contract=off",
                           "-Wno-unknown-pragmas", "-DVIDO_EMUL_CHECK_ALIGN", "-o", exe, os.path.join(ROOT, "tests", "desc_emul_asan.cc")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ASAN_WALK_OK" in out.stdout
