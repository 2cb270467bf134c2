"""Golden cases for the initial camera model (Tracking::GetInitModelCam, src/Tracking.cc:1914-2028) from the OpenCV call the
reference makes: cv::solvePnPRansac(pts3d, pts2d, K, noArray, rvec, tvec, false, 500, 0.4, 0.98, inliers, SOLVEPNP_P3P)
(src/Tracking.cc:1967) followed by cv::Rodrigues, and the constant-velocity alternative scored with the same 0.4 px gate
(:1980-2001).  Run in the authoring container (cv2 4.13; the reference linked OpenCV 3.4 -- SURVEY.md 8c'):

    python tests/golden/make_pnp_golden.py      -> tests/golden/pnp_golden.npz

The cases include constant-velocity priors that are badly wrong (the product's minimal solver is seeded from the
prior, so RANSAC must still find the pose when the prior is poor)."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pose_synth  # noqa: E402

CASES = [  # n, seed, outliers, motion_err (metres along the optical axis, mostly)
    (1200, 1, 0.2, 0.3), (1200, 2, 0.5, 1.0), (2400, 3, 0.05, 0.0), (800, 4, 0.7, 0.5), (40, 5, 0.1, 0.2),
    (1000, 6, 0.3, 3.0), (1000, 7, 0.3, -4.0), (600, 8, 0.6, 8.0),   # poor priors
]


def motion_model_inliers(pts, cur, T, K, thr=0.4):
    """float32 arithmetic of the reference's loop (cv::Mat CV_32F products), src/Tracking.cc:1980-2001"""
    fx, fy, cx, cy = [np.float32(v) for v in K]
    T = T.astype(np.float32)
    ids = []
    for i in range(len(pts)):
        x3 = pts[i].astype(np.float32)
        pc = np.array([np.float32(np.float32(np.float32(np.float32(T[r, 0] * x3[0]) + np.float32(T[r, 1] * x3[1])) + np.float32(T[r, 2] * x3[2])) + T[r, 3])
                       for r in range(3)], np.float32)
        invz = np.float32(1.0 / np.float64(pc[2]))
        u = np.float32(np.float32(fx * pc[0]) * invz + cx)
        v = np.float32(np.float32(fy * pc[1]) * invz + cy)
        du, dv = np.float32(cur[i, 0] - u), np.float32(cur[i, 1] - v)
        if np.sqrt(np.float32(du * du + dv * dv)) < np.float32(thr):
            ids.append(i)
    return np.array(ids, np.int32)


def main():
    cv2.setNumThreads(1)
    out = {"cv2_version": np.array(cv2.__version__)}
    for k, (n, seed, outl, merr) in enumerate(CASES):
        pr = pose_synth.make_pnp(n=n, seed=seed, outliers=outl, motion_err=merr)
        fx, fy, cx, cy = pr["K"]
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
        ok, rvec, tvec, inl = cv2.solvePnPRansac(pr["pts"].reshape(-1, 1, 3), pr["cur"].reshape(-1, 1, 2), K, None, iterationsCount=500,
                                                 reprojectionError=0.4, confidence=0.98, flags=cv2.SOLVEPNP_P3P)
        T = np.eye(4)
        if ok:
            T[:3, :3] = cv2.Rodrigues(rvec)[0]
            T[:3, 3] = tvec.reshape(3)
        inl = np.zeros(0, np.int32) if inl is None else inl.reshape(-1).astype(np.int32)
        mm = motion_model_inliers(pr["pts"], pr["cur"], pr["Tcw_motion"], pr["K"])
        winner = 0 if len(inl) > len(mm) else 1     # src/Tracking.cc:2006: the RANSAC model wins with MORE inliers
        out[f"c{k}_params"] = np.array([n, seed, outl, merr], np.float64)
        out[f"c{k}_cur"] = pr["cur"]; out[f"c{k}_pts"] = pr["pts"]; out[f"c{k}_Tmm"] = pr["Tcw_motion"]
        out[f"c{k}_K"] = np.array(pr["K"], np.float64); out[f"c{k}_Tgt"] = pr["Tcw_gt"]
        out[f"c{k}_cv_T"] = T.astype(np.float32); out[f"c{k}_cv_inl"] = inl; out[f"c{k}_mm_inl"] = mm
        out[f"c{k}_winner"] = np.array(winner, np.int32); out[f"c{k}_ok"] = np.array(int(bool(ok)), np.int32)
        print(f"case {k}: n={n} outliers={outl} motion_err={merr}: cv2 ok={ok} inliers={len(inl)} motion-model inliers={len(mm)} winner={winner}")
    out["n_cases"] = np.array(len(CASES), np.int32)
    np.savez_compressed(os.path.join(HERE, "pnp_golden.npz"), **out)


if __name__ == "__main__":
    main()
