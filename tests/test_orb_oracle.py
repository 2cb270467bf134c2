"""CPU: the oracle's ORB restatement against the cv2-generated golden vectors (tests/golden/orb_golden.npz)."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLD = os.path.join(os.path.dirname(__file__), "golden", "orb_golden.npz")
CASES = ["kitti_scene", "small_scene", "noise_333x211"]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_level_sizes_and_quotas():
    p = ol.default_orb_params()
    w, h, s = ol.level_sizes(1242, 375, p)
    assert list(zip(w.tolist(), h.tolist())) == [(1242, 375), (1035, 312), (862, 260), (719, 217), (599, 181),
                                                  (499, 151), (416, 126), (347, 105)]
    assert ol.level_quotas(p).tolist() == [543, 452, 377, 314, 262, 218, 182, 152]
    assert ol.level_quotas(p).sum() == 2500


@pytest.mark.parametrize("name", CASES)
def test_pyramid_bytes_match_cv2(gold, name):
    p = ol.default_orb_params()
    pyr = ol.orb_pyramid(gold[f"{name}_img"], p)
    sizes = gold[f"{name}_sizes"]
    for l, lvl in enumerate(pyr):
        assert lvl.shape == (sizes[l, 1], sizes[l, 0])
        assert hashlib.sha256(lvl.tobytes()).hexdigest() == str(gold[f"{name}_pyr_sha"][l]), f"level {l}"


@pytest.mark.parametrize("name", CASES)
def test_fast_candidates_match_cv2(gold, name):
    """per-cell cv::FAST (thr 20, fallback 7, cell-local NMS) list, order and scores, every level"""
    p = ol.default_orb_params()
    pyr = ol.orb_pyramid(gold[f"{name}_img"], p)
    checked = 0
    for l, lvl in enumerate(pyr):
        key = f"{name}_cand{l}"
        if key not in gold:
            continue
        xs, ys, sc = ol.level_candidates(lvl, p)
        ref = gold[key]
        assert len(xs) == len(ref), f"level {l}"
        assert np.array_equal(np.stack([xs, ys, sc], 1), ref), f"level {l}"
        checked += 1
    assert checked >= 6


def test_fast_atan2_matches_cv2(gold):
    yx = gold["atan2_yx"]
    mine = np.array([ol.fast_atan2(a, b) for a, b in yx], np.float32)
    assert np.array_equal(mine.view(np.uint32), gold["atan2_ref"].view(np.uint32))


@pytest.mark.parametrize("name", CASES)
def test_keypoints_regression(gold, name):
    """full extractor output is stable (quad-tree restatement is the spec; regression vector)"""
    p = ol.default_orb_params()
    kp = ol.orb_extract(gold[f"{name}_img"], p)
    ref = gold[f"{name}_oracle_kp"]
    assert kp.tobytes() == ref.tobytes()


def test_keypoint_invariants(gold):
    p = ol.default_orb_params()
    img = gold["kitti_scene_img"]
    kp = ol.orb_extract(img, p)
    H, W = img.shape
    q = ol.level_quotas(p)
    cnt = np.bincount(kp["octave"], minlength=8)
    assert np.all(cnt <= q + 3) and np.all(np.diff(kp["octave"]) >= 0)
    assert kp["x"].min() >= 19 and kp["x"].max() < W - 19 + 1.2 ** 7
    assert kp["y"].min() >= 19 and kp["y"].max() < H
    assert np.all((kp["angle"] >= 0) & (kp["angle"] < 360))
    assert np.all(kp["response"] >= 7)
    # no duplicates within a level
    for l in range(8):
        k = kp[kp["octave"] == l]
        assert len(np.unique(np.stack([k["x"], k["y"]], 1), axis=0)) == len(k)


def test_empty_and_flat_images():
    p = ol.default_orb_params()
    flat = np.full((375, 1242), 128, np.uint8)
    assert len(ol.orb_extract(flat, p)) == 0
    xs, _, _ = ol.fast_roi(np.zeros((6, 6), np.uint8), 7)
    assert len(xs) == 0
