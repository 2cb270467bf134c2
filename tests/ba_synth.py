"""Synthetic sliding-window BA problems in the reference's Map conventions (float32 Twc poses, world points)."""
import numpy as np


def rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def make_window(W=20, P=600, seed=0, obs_noise=0.02, pose_noise=0.01, rot_noise=0.002, outliers=0.02, min_len=3):
    """Returns dict of float32 arrays: poses [W,16], rel [W-1,16], points [P,3], obs_pose, obs_point, obs_xyz
    and ground truth.  Tracks are born at a random window frame and last >= min_len frames (as in
    PartialBatchOptimization: only tracks born inside the window, length >= 3)."""
    rng = np.random.default_rng(seed)
    gt = []
    for k in range(W):
        T = np.eye(4)
        T[:3, :3] = rot([0, 1, 0], 0.03 * np.sin(0.3 * k)) @ rot([1, 0, 0], 0.004 * np.cos(0.2 * k))
        T[:3, 3] = [0.5 * np.sin(0.1 * k), 0.02 * np.sin(0.4 * k), 1.0 * k]
        gt.append(T)
    gt = np.array(gt)
    est = gt.copy()
    for k in range(1, W):
        dT = np.eye(4)
        dT[:3, :3] = rot(rng.normal(size=3), rot_noise * rng.normal())
        dT[:3, 3] = pose_noise * rng.normal(size=3)
        est[k] = est[k] @ dT
    obs_pose, obs_point, obs_xyz, pts = [], [], [], []
    for l in range(P):
        born = int(rng.integers(0, max(W - min_len + 1, 1)))
        length = int(rng.integers(min_len, W - born + 1)) if W - born >= min_len else W - born
        # a point in front of the camera at its birth frame, 4..40 m away
        pc = np.array([rng.uniform(-8, 8), rng.uniform(-3, 1.5), rng.uniform(4 + length, 40 + length)])
        pw = gt[born][:3, :3] @ pc + gt[born][:3, 3]
        # initial estimate = back-projection through the (noisy) estimated pose of the birth frame
        z0 = pc + obs_noise * rng.normal(size=3) * (pc[2] / 10)
        pts.append(est[born][:3, :3] @ z0 + est[born][:3, 3])
        for k in range(born, born + length):
            Tcw = np.linalg.inv(gt[k])
            z = Tcw[:3, :3] @ pw + Tcw[:3, 3]
            z = z + obs_noise * rng.normal(size=3) * (z[2] / 10)
            if rng.uniform() < outliers:
                z = z + rng.normal(size=3) * 2.0
            if k == born:
                z = z0
            obs_pose.append(k); obs_point.append(l); obs_xyz.append(z)
    # observations are stored frame-major like the reference's graph construction
    order = np.lexsort((np.array(obs_point), np.array(obs_pose)))
    poses32 = est.reshape(W, 16).astype(np.float32)
    rel = np.array([np.linalg.inv(est[k - 1]) @ est[k] for k in range(1, W)]).reshape(W - 1, 16)
    # the odometry measurement comes from the per-frame tracker: perturb it independently
    rel = rel + 0.0
    return dict(poses=poses32, rel=rel.astype(np.float32), points=np.array(pts, np.float32),
                obs_pose=np.array(obs_pose, np.int32)[order], obs_point=np.array(obs_point, np.int32)[order],
                obs_xyz=np.array(obs_xyz, np.float32)[order], gt=gt)
