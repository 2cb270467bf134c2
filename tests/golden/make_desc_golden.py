"""Generates tests/golden/desc_golden.npz with cv2 (OpenCV is the un-vendored owner of these stages; the reference links 3.4).
Run in the authoring container:   python tests/golden/make_desc_golden.py
Pins, per case image:
  * blur{l}_sha   sha256 of cv2.GaussianBlur(level, (7,7), 2, 2, BORDER_REFLECT_101) for every pyramid level
                  (src/ORBextractor.cc:1078-1079), the level being the cv2.resize pyramid orb_golden.npz pins;
  * desc_cv       cv2.ORB.compute descriptors of the oracle's key points, level by level: each level is handed to cv2 as an image
                  of its own with the key points at their level coordinates, octave 0 and the oracle's angle, so that cv2's own
                  pyramid / orientation code is not involved -- what is compared is computeOrbDescriptors on the blurred level
                  (the reference's computeOrbDescriptor, src/ORBextractor.cc:98-137, is OpenCV's function).  cv2.ORB smooths its
                  pyramid as a sub-matrix in place, and OpenCV sends an 8-bit SUB-MATRIX through the float-kernel separable filter
                  (result = the correctly rounded float Gaussian) instead of the fixed-point smoother a stand-alone Mat -- the
                  reference's clone -- gets; the two differ by one grey level on a fifth of the pixels.  So desc_cv pins the
                  descriptor function on cv2's own smoothing: sepblur{l}_sha = sha256 of cv2.sepFilter2D(level, getGaussianKernel(7, 2)),
                  which tests/test_desc_oracle.py reproduces with a float64 numpy model (plus sepblur_fix: the few pixels where
                  cv2's float32 accumulation lands on the other side of a x.5 tie) and feeds to the oracle's describe function;
                  cv2 drops key points closer than its edgeThreshold to the border: desc_idx lists the rows it kept
                  (edgeThreshold 19 = the distance up to which every rotated test point stays inside the level);
  * match_*       cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) between the descriptor sets of two frames.
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


def cv_pyramid(img, p):
    H, W = img.shape
    w, h, s = ol.level_sizes(W, H, p)
    pyr = [img]
    for l in range(1, p.nlevels):
        pyr.append(cv2.resize(pyr[-1], (int(w[l]), int(h[l])), interpolation=cv2.INTER_LINEAR))
    return pyr, s


def sep_blur_model(img):
    """float64 numpy model of OpenCV's separable-filter smoothing of an 8-bit image (same function in tests/test_desc_oracle.py)"""
    k = np.exp(-np.arange(-3, 4) ** 2 / 8.0)
    k /= k.sum()
    H, W = img.shape
    pad = np.pad(img.astype(np.float64), 3, mode="reflect")
    h = sum(k[j] * pad[:, j:j + W] for j in range(7))
    v = sum(k[j] * h[j:j + H, :] for j in range(7))
    return np.rint(v).clip(0, 255).astype(np.uint8)


def cv_descriptors(pyr, scale, kps):
    """descriptors by cv2.ORB.compute, one level at a time; returns (row indices kept, descriptors)"""
    orb = cv2.ORB_create(nfeatures=100000, scaleFactor=1.2, nlevels=1, edgeThreshold=19, firstLevel=0, WTA_K=2, patchSize=31)
    rows, descs = [], []
    for l, lvl in enumerate(pyr):
        sel = np.nonzero(kps["octave"] == l)[0]
        if len(sel) == 0:
            continue
        cvk = []
        for i in sel:
            x = float(np.rint(np.float32(kps["x"][i]) / np.float32(scale[l]))) if l else float(kps["x"][i])
            y = float(np.rint(np.float32(kps["y"][i]) / np.float32(scale[l]))) if l else float(kps["y"][i])
            k = cv2.KeyPoint(x, y, 31.0, float(kps["angle"][i]), float(kps["response"][i]), 0, int(i))
            cvk.append(k)
        kept, d = orb.compute(lvl, cvk)
        for k, row in zip(kept, d):
            rows.append(k.class_id)
            descs.append(row)
    order = np.argsort(rows)
    return np.array(rows, np.int32)[order], np.array(descs, np.uint8)[order]


def main():
    p = ol.default_orb_params()
    data = {"cv2_version": np.array(cv2.__version__)}
    scene = synth.Scene(cam=synth.KITTI, seed=1234)
    cases = {
        "kitti_scene": scene.frame(3)["gray"].numpy(),
        "kitti_scene_next": scene.frame(4)["gray"].numpy(),
        "noise_333x211": synth.noise_image(333, 211, 7),
    }
    descs = {}
    for name, img in cases.items():
        pyr, scale = cv_pyramid(img, p)
        data[f"{name}_img"] = img
        data[f"{name}_blur_sha"] = np.array([hashlib.sha256(cv2.GaussianBlur(l, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101).tobytes()).hexdigest() for l in pyr])
        k32 = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
        sep = [cv2.sepFilter2D(l, -1, k32, k32, borderType=cv2.BORDER_REFLECT_101) for l in pyr]
        fix = []   # (level, y, x, value) where cv2's float32 accumulation resolves a x.5 tie differently from the float64 model
        for li, (l, sb) in enumerate(zip(pyr, sep)):
            m = sep_blur_model(l)
            ys, xs = np.nonzero(m != sb)
            assert len(ys) < 1e-5 * m.size + 4 and (np.abs(m.astype(int) - sb)[ys, xs] == 1).all()
            fix += [(li, y, x, sb[y, x]) for y, x in zip(ys, xs)]
        data[f"{name}_sepblur_fix"] = np.array(fix, np.int32).reshape(-1, 4)
        data[f"{name}_sepblur_sha"] = np.array([hashlib.sha256(sb.tobytes()).hexdigest() for sb in sep])
        kps = ol.orb_extract(img, p)
        idx, d = cv_descriptors(pyr, scale, kps)
        data[f"{name}_oracle_kp"] = kps
        data[f"{name}_desc_idx"] = idx
        data[f"{name}_desc_cv"] = d
        descs[name] = d
        print(name, len(kps), "key points,", len(idx), "described by cv2")
    q, t = descs["kitti_scene"], descs["kitti_scene_next"]
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    data["match_best_idx"] = np.array([m[0].trainIdx for m in knn], np.int32)
    data["match_best_dist"] = np.array([int(m[0].distance) for m in knn], np.int32)
    data["match_second_dist"] = np.array([int(m[1].distance) for m in knn], np.int32)
    one = cv2.BFMatcher(cv2.NORM_HAMMING).match(q, t)
    data["match1_best_idx"] = np.array([m.trainIdx for m in one], np.int32)
    np.savez_compressed(os.path.join(HERE, "desc_golden.npz"), **data)
    print("wrote desc_golden.npz")


if __name__ == "__main__":
    main()
