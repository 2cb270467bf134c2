"""Small hand-made FullBatch graphs (Optimizer::FullBatchOptimization, src/Optimizer.cc:1235-1745) with known ground truth:
a camera driving forward, static points seen over several frames, rigid objects whose points are re-observed every frame
(one point vertex per observation, tied by the object's world-frame motion)."""
import numpy as np


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _T(R, t):
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return T


def make_graph(n_frames=6, n_static=60, n_objects=2, pts_per_obj=12, seed=0, noise=0.01, track_len=4):
    """returns (g dict keyed like FBA_KEYS, n_poses, truth dict)"""
    rng = np.random.default_rng(seed)
    Twc = [_T(_rot(0.01 * np.sin(k), 0.03 * np.sin(0.7 * k), 0.0), [0.2 * np.sin(0.5 * k), 0.0, 1.0 * k]) for k in range(n_frames)]
    Twc = [np.linalg.inv(Twc[0]) @ T for T in Twc]
    Hobj = [[_T(_rot(0, 0.02 * (j + 1), 0), [0.05 * j, 0.0, 0.8 + 0.1 * j]) for _ in range(n_frames)] for j in range(n_objects)]
    se3 = [T.copy() for T in Twc]
    points, obs_se3, obs_point, obs_kind, obs_xyz = [], [], [], [], []
    e6_i, e6_j, e6_kind, e6_meas = [], [], [], []
    tern_p1, tern_p2, tern_h = [], [], []
    truth_pts = []
    for k in range(1, n_frames):
        e6_i.append(k - 1); e6_j.append(k); e6_kind.append(0)
        Z = np.linalg.inv(Twc[k - 1]) @ Twc[k]
        Z[:3, 3] += noise * 0.2 * rng.standard_normal(3)
        e6_meas.append(Z)
    # static points: each seen in `track_len` consecutive frames starting at a random frame
    for _ in range(n_static):
        f0 = int(rng.integers(0, max(n_frames - track_len, 0) + 1))
        Xw = np.array([rng.uniform(-6, 6), rng.uniform(-2, 1.5), f0 + rng.uniform(6, 25)])
        pid = len(points)
        truth_pts.append(Xw)
        points.append(Xw + noise * 5 * rng.standard_normal(3))
        for k in range(f0, min(f0 + track_len, n_frames)):
            xc = np.linalg.inv(Twc[k])[:3, :3] @ Xw + np.linalg.inv(Twc[k])[:3, 3]
            obs_se3.append(k); obs_point.append(pid); obs_kind.append(0); obs_xyz.append(xc + noise * rng.standard_normal(3))
    # objects: motion vertex per frame >= 1, smoothness from frame 3 on, a chain of point vertices per object point
    motion_vid = {}
    for k in range(1, n_frames):
        for j in range(n_objects):
            vid = len(se3)
            se3.append(np.eye(4))
            motion_vid[(k, j)] = vid
            if k > 2:
                e6_i.append(motion_vid[(k - 1, j)]); e6_j.append(vid); e6_kind.append(1); e6_meas.append(np.eye(4))
    for j in range(n_objects):
        for _ in range(pts_per_obj):
            Xw = np.array([rng.uniform(-1, 1) + 3 * j - 2, rng.uniform(-0.5, 0.5), rng.uniform(9, 11)])
            prev = -1
            for k in range(n_frames):
                if k > 0:
                    Xw = Hobj[j][k][:3, :3] @ Xw + Hobj[j][k][:3, 3]
                pid = len(points)
                truth_pts.append(Xw.copy())
                points.append(Xw + noise * 5 * rng.standard_normal(3))
                xc = np.linalg.inv(Twc[k])[:3, :3] @ Xw + np.linalg.inv(Twc[k])[:3, 3]
                obs_se3.append(k); obs_point.append(pid); obs_kind.append(1); obs_xyz.append(xc + noise * rng.standard_normal(3))
                if k > 0:
                    tern_p1.append(prev); tern_p2.append(pid); tern_h.append(motion_vid[(k, j)])
                prev = pid
    # perturbed initial poses (frame 0 stays: it carries the prior)
    for k in range(1, n_frames):
        se3[k] = se3[k] @ _T(_rot(*(0.002 * rng.standard_normal(3))), 0.03 * rng.standard_normal(3))
    g = dict(se3=np.array([T.reshape(16) for T in se3], np.float32), points=np.array(points, np.float32),
             e6_i=np.array(e6_i, np.int32), e6_j=np.array(e6_j, np.int32), e6_kind=np.array(e6_kind, np.int32),
             e6_meas=np.array([Z.reshape(16) for Z in e6_meas], np.float32).reshape(-1, 16),
             obs_se3=np.array(obs_se3, np.int32), obs_point=np.array(obs_point, np.int32), obs_kind=np.array(obs_kind, np.int32),
             obs_xyz=np.array(obs_xyz, np.float32), tern_p1=np.array(tern_p1, np.int32), tern_p2=np.array(tern_p2, np.int32),
             tern_h=np.array(tern_h, np.int32))
    truth = dict(Twc=np.array(Twc), H=np.array(Hobj), points=np.array(truth_pts), motion_vid=motion_vid)
    return g, n_frames, truth
