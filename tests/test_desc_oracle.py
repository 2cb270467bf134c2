"""CPU: the oracle's descriptor stage (7x7 Gaussian, rBRIEF, Hamming matching) against the cv2-generated golden vectors
(tests/golden/desc_golden.npz, made by tests/golden/make_desc_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLD = os.path.join(os.path.dirname(__file__), "golden", "desc_golden.npz")
CASES = ["kitti_scene", "kitti_scene_next", "noise_333x211"]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("name", CASES)
def test_blurred_levels_match_cv2(gold, name):
    """GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of every pyramid level, byte for byte (src/ORBextractor.cc:1078-1079)"""
    p = ol.default_orb_params()
    pyr = ol.orb_pyramid(gold[f"{name}_img"], p)
    for l, lvl in enumerate(pyr):
        assert hashlib.sha256(ol.gauss7(lvl).tobytes()).hexdigest() == str(gold[f"{name}_blur_sha"][l]), f"level {l}"


def sep_blur_model(img):
    """float64 numpy model of OpenCV's separable-filter smoothing of an 8-bit image: what cv2.ORB applies to its pyramid (a
    sub-matrix, smoothed in place, does not take the fixed-point path); equal to cv2.sepFilter2D byte for byte (sepblur_sha)"""
    k = np.exp(-np.arange(-3, 4) ** 2 / 8.0)
    k /= k.sum()
    H, W = img.shape
    pad = np.pad(img.astype(np.float64), 3, mode="reflect")
    h = sum(k[j] * pad[:, j:j + W] for j in range(7))
    v = sum(k[j] * h[j:j + H, :] for j in range(7))
    return np.rint(v).clip(0, 255).astype(np.uint8)


@pytest.mark.parametrize("name", CASES)
def test_descriptor_function_matches_cv2(gold, name):
    """computeOrbDescriptor (src/ORBextractor.cc:98-137) against cv2.ORB.compute on the same smoothed level, key point and angle:
    every bit of every descriptor, all levels"""
    p = ol.default_orb_params()
    img = gold[f"{name}_img"]
    kps = gold[f"{name}_oracle_kp"]
    idx = gold[f"{name}_desc_idx"]
    assert np.array_equal(idx, np.arange(len(kps)))                 # cv2 kept every key point (all are >= 19 px from the border)
    _, _, s = ol.level_sizes(img.shape[1], img.shape[0], p)
    checked = 0
    for l, lvl in enumerate(ol.orb_pyramid(img, p)):
        sel = np.nonzero(kps["octave"] == l)[0]
        if len(sel) == 0:
            continue
        sb = sep_blur_model(lvl)
        for fl, fy, fx, fv in gold[f"{name}_sepblur_fix"]:           # x.5 ties cv2's float32 sums break the other way (3 pixels)
            if fl == l:
                sb[fy, fx] = fv
        assert hashlib.sha256(sb.tobytes()).hexdigest() == str(gold[f"{name}_sepblur_sha"][l])
        lx = np.rint(kps["x"][sel] / s[l]) if l else kps["x"][sel]
        ly = np.rint(kps["y"][sel] / s[l]) if l else kps["y"][sel]
        d = ol.describe_level(sb, lx, ly, kps["angle"][sel])
        assert np.array_equal(d, gold[f"{name}_desc_cv"][sel]), f"level {l}"
        checked += len(sel)
    assert checked == len(kps)


@pytest.mark.parametrize("name", CASES)
def test_extract_describe_is_blur_then_describe(gold, name):
    """operator() with the descriptor call enabled = the pinned smoother (the reference's clone: fixed-point path) followed by the
    pinned descriptor function at the LEVEL coordinates; the key points themselves are unchanged"""
    p = ol.default_orb_params()
    img = gold[f"{name}_img"]
    kps, desc = ol.orb_extract_describe(img, p)
    assert np.array_equal(kps, gold[f"{name}_oracle_kp"])
    assert np.array_equal(kps, ol.orb_extract(img, p))
    _, _, s = ol.level_sizes(img.shape[1], img.shape[0], p)
    for l, lvl in enumerate(ol.orb_pyramid(img, p)):
        sel = np.nonzero(kps["octave"] == l)[0]
        if len(sel) == 0:
            continue
        lx = np.rint(kps["x"][sel] / s[l]) if l else kps["x"][sel]
        ly = np.rint(kps["y"][sel] / s[l]) if l else kps["y"][sel]
        assert np.array_equal(desc[sel], ol.describe_level(ol.gauss7(lvl), lx, ly, kps["angle"][sel])), f"level {l}"
    assert len(np.unique(desc, axis=0)) > 0.95 * len(desc)          # not degenerate
    # the two smoothers differ by a grey level here and there: a handful of bits per descriptor, never the descriptor's identity
    bits = np.unpackbits(desc ^ gold[f"{name}_desc_cv"], axis=1).sum(1)
    assert bits.max() <= 12 and bits.mean() < 1.5


def test_key_points_stay_19_pixels_inside(gold):
    """FAST runs on cells that start at minBorder = 16 and detects from 3 px inside a cell: every key point is >= 19 px from the
    level border, the reach of a rotated test point (|(13, 13)| = 18.4) -- the descriptor never reads outside the blurred level"""
    p = ol.default_orb_params()
    for name in CASES:
        img = gold[f"{name}_img"]
        w, h, s = ol.level_sizes(img.shape[1], img.shape[0], p)
        k = gold[f"{name}_oracle_kp"]
        lx = np.rint(k["x"] / s[k["octave"]]); ly = np.rint(k["y"] / s[k["octave"]])
        assert (lx >= 19).all() and (ly >= 19).all()
        assert (lx <= w[k["octave"]] - 20).all() and (ly <= h[k["octave"]] - 20).all()


def test_describe_level_border_rule():
    """a key point whose rotated pattern leaves the buffer reads 0 there; horizontal overshoot wraps into the neighbouring row"""
    rng = np.random.default_rng(3)
    img = rng.integers(1, 256, (64, 80), dtype=np.uint8)
    d_in = ol.describe_level(img, [40.0], [32.0], [37.0])
    # same key point in an image padded by rows of zeros above and below: identical reads, so identical descriptor
    pad = np.zeros((64 + 40, 80), np.uint8); pad[20:84] = img
    assert np.array_equal(ol.describe_level(pad, [40.0], [52.0], [37.0]), d_in)
    # at the top edge the out-of-buffer reads are zeros = the padded image's zero rows
    d_top = ol.describe_level(img, [40.0], [5.0], [37.0])
    assert np.array_equal(ol.describe_level(pad, [40.0], [25.0], [37.0]), d_top)


def test_hamming_match_matches_cv2(gold):
    q, t = gold["kitti_scene_desc_cv"], gold["kitti_scene_next_desc_cv"]   # 2513 x 2516 descriptors of two consecutive frames
    bi, bd, sd = ol.hamming_match(q, t)
    assert np.array_equal(bd, gold["match_best_dist"])
    assert np.array_equal(sd, gold["match_second_dist"])
    assert np.array_equal(bi, gold["match1_best_idx"])               # BFMatcher.match: the first minimum
    # knnMatch orders equal distances its own way: wherever best < second the index is unambiguous
    clear = bd < sd
    assert np.array_equal(bi[clear], gold["match_best_idx"][clear])
    assert (bd[np.arange(len(q))] == [int(np.unpackbits(q[i] ^ t[bi[i]]).sum()) for i in range(len(q))]).all()


def test_hamming_match_small_sets():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 256, (5, 32), dtype=np.uint8)
    bi, bd, sd = ol.hamming_match(q, np.zeros((0, 32), np.uint8))
    assert (bi == -1).all() and (bd == 0x7fffffff).all() and (sd == 0x7fffffff).all()
    bi, bd, sd = ol.hamming_match(q, q[:1])
    assert (bi == 0).all() and bd[0] == 0 and (sd == 0x7fffffff).all()
    bi, bd, sd = ol.hamming_match(q, np.concatenate([q, q]))         # duplicates: lowest index wins, second distance = 0
    assert np.array_equal(bi, np.arange(5)) and (bd == 0).all() and (sd == 0).all()
