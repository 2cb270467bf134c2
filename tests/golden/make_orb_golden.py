"""Generates tests/golden/orb_golden.npz with cv2 (the only faithful implementation of the OpenCV-resident
stages available here: the reference links OpenCV 3.4, un-vendored).  Run in the authoring container:
    python tests/golden/make_orb_golden.py
Pins: cv::resize INTER_LINEAR 8U pyramid bytes, per-cell cv::FAST(thr 20 -> 7, NMS) candidate lists in the
reference's order (src/ORBextractor.cc:779-819), cv::fastAtan2 samples.  The quad-tree has no external
implementation; its expected output is produced by the oracle restatement and stored as a regression vector
(flagged 'oracle_*').
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import synth  # noqa: E402

cv2.setNumThreads(1)


def cv_level_candidates(img, ini=20, mn=7):
    rows, cols = img.shape
    minB = 16
    maxBX, maxBY = cols - 16, rows - 16
    width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
    nCols, nRows = int(width / np.float32(30)), int(height / np.float32(30))
    wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
    d20 = cv2.FastFeatureDetector_create(threshold=ini, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    d7 = cv2.FastFeatureDetector_create(threshold=mn, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    out = []
    for i in range(nRows):
        iniY = minB + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minB + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            roi = np.ascontiguousarray(img[iniY:maxY, iniX:maxX])
            kps = d20.detect(roi, None)
            if len(kps) == 0:
                kps = d7.detect(roi, None)
            for k in kps:
                out.append((int(k.pt[0]) + j * wCell, int(k.pt[1]) + i * hCell, int(k.response)))
    return np.array(out, np.int32).reshape(-1, 3)


def main():
    p = ol.default_orb_params()
    data = {"cv2_version": np.array(cv2.__version__), "cv2_build": np.array(cv2.getBuildInformation()[:2000])}
    cases = {
        "kitti_scene": synth.Scene(cam=synth.KITTI, seed=1234).frame(3)["gray"].numpy(),
        "small_scene": synth.Scene(cam=synth.SMALL, seed=1235).frame(1)["gray"].numpy(),
        "noise_333x211": synth.noise_image(333, 211, 7),
    }
    for name, img in cases.items():
        H, W = img.shape
        w, h, _ = ol.level_sizes(W, H, p)
        cur = img
        digests = []
        for l in range(p.nlevels):
            if l > 0:
                cur = cv2.resize(cur, (int(w[l]), int(h[l])), interpolation=cv2.INTER_LINEAR)
            digests.append(hashlib.sha256(cur.tobytes()).hexdigest())
            if min(cur.shape) > 70:
                data[f"{name}_cand{l}"] = cv_level_candidates(cur)
        data[f"{name}_img"] = img
        data[f"{name}_pyr_sha"] = np.array(digests)
        data[f"{name}_sizes"] = np.stack([w, h], 1)
        kp = ol.orb_extract(img, p)
        data[f"{name}_oracle_kp"] = kp
    rng = np.random.default_rng(5)
    yx = rng.integers(-60000, 60000, (4000, 2)).astype(np.float32)
    yx[:8] = [[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0], [5, 5], [-5, 5], [5, -5]]
    data["atan2_yx"] = yx
    data["atan2_ref"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "orb_golden.npz"), **data)
    print("wrote orb_golden.npz", {k: getattr(v, "shape", None) for k, v in data.items()})


if __name__ == "__main__":
    main()
