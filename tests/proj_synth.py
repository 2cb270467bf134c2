"""Synthetic 3-D / 2-D correspondences for the reprojection-only optimisers (PoseOptimizationNew / PoseOptimizationObjMot)."""
import numpy as np

import fba_synth

K = (718.856, 718.856, 607.1928, 185.2157)


def camera_case(n=400, seed=0, outliers=0.1, noise=0.05):
    """kind 0: world points, true camera pose Tcw, observations = projection + noise; a perturbed initial pose"""
    rng = np.random.default_rng(seed)
    Tcw = fba_synth._T(fba_synth._rot(0.01, -0.03, 0.005), [0.1, -0.05, -1.0])
    X = np.stack([rng.uniform(-8, 8, n), rng.uniform(-2, 1.5, n), rng.uniform(6, 40, n)], 1)
    pc = X @ Tcw[:3, :3].T + Tcw[:3, 3]
    uv = np.stack([K[0] * pc[:, 0] / pc[:, 2] + K[2], K[1] * pc[:, 1] / pc[:, 2] + K[3]], 1) + noise * rng.standard_normal((n, 2))
    bad = rng.random(n) < outliers
    uv[bad] += rng.uniform(3, 30, (int(bad.sum()), 2))
    T0 = fba_synth._T(fba_synth._rot(0.004, 0.006, -0.003), [0.05, 0.02, -0.04]) @ Tcw
    return dict(kind=0, obs_xy=uv.astype(np.float32), pts3d=X.astype(np.float32), T_init=T0.astype(np.float32), K=K), Tcw, bad


def object_case(n=300, seed=0, outliers=0.1, noise=0.05):
    """kind 1: points on an object in the world at the last frame, true world-frame motion H, camera pose Tcw; P = K * Tcw"""
    rng = np.random.default_rng(seed)
    Tcw = fba_synth._T(fba_synth._rot(0.0, 0.02, 0.0), [0.2, 0.0, -3.0])
    H = fba_synth._T(fba_synth._rot(0.0, 0.03, 0.0), [0.1, 0.0, 0.9])
    X = np.stack([rng.uniform(-1, 1, n) + 2.0, rng.uniform(-0.7, 0.7, n), rng.uniform(11, 13, n)], 1)
    Xn = X @ H[:3, :3].T + H[:3, 3]
    pc = Xn @ Tcw[:3, :3].T + Tcw[:3, 3]
    uv = np.stack([K[0] * pc[:, 0] / pc[:, 2] + K[2], K[1] * pc[:, 1] / pc[:, 2] + K[3]], 1) + noise * rng.standard_normal((n, 2))
    bad = rng.random(n) < outliers
    uv[bad] += rng.uniform(3, 30, (int(bad.sum()), 2))
    Kp = np.array([[K[0], 0, K[2], 0], [0, K[1], K[3], 0], [0, 0, 1, 0]], np.float64)
    P = Kp @ Tcw.astype(np.float32).astype(np.float64)
    H0 = fba_synth._T(fba_synth._rot(0.0, 0.005, 0.0), [0.0, 0.0, 0.1]) @ H
    return dict(kind=1, obs_xy=uv.astype(np.float32), pts3d=X.astype(np.float32), T_init=H0.astype(np.float32), P=P), H, bad
