"""Synthetic inputs of the per-frame pose optimisation (PoseOptimizationFlow2Cam)."""
import numpy as np

import ba_synth


def make_poseopt(n=800, seed=0, flow_noise=0.1, outliers=0.05, cam=(718.856, 718.856, 607.1928, 185.2157),
                 W=1242, H=375, init_noise=0.02):
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = cam
    Twl = np.eye(4)
    Twl[:3, :3] = ba_synth.rot([0, 1, 0], 0.1)
    Twl[:3, 3] = [0.3, 0.0, 5.0]
    dT = np.eye(4)
    dT[:3, :3] = ba_synth.rot([0.1, 1, 0.05], 0.02)
    dT[:3, 3] = [0.02, -0.01, 1.0]
    Twc = Twl @ dT
    Tcw, Tlw = np.linalg.inv(Twc), np.linalg.inv(Twl)
    obs = np.stack([rng.uniform(20, W - 20, n), rng.uniform(20, H - 20, n)], 1)
    depth = rng.uniform(4, 40, n)
    Xc = np.stack([(obs[:, 0] - cx) * depth / fx, (obs[:, 1] - cy) * depth / fy, depth], 1)
    Xw = Xc @ Twl[:3, :3].T + Twl[:3, 3]
    Xn = Xw @ Tcw[:3, :3].T + Tcw[:3, 3]
    proj = np.stack([fx * Xn[:, 0] / Xn[:, 2] + cx, fy * Xn[:, 1] / Xn[:, 2] + cy], 1)
    flow = proj - obs + flow_noise * rng.normal(size=(n, 2))
    bad = rng.uniform(size=n) < outliers
    flow[bad] += rng.normal(size=(bad.sum(), 2)) * 8
    init = np.eye(4)
    init[:3, :3] = ba_synth.rot(rng.normal(size=3), init_noise * 0.1) @ Tcw[:3, :3]
    init[:3, 3] = Tcw[:3, 3] + init_noise * rng.normal(size=3)
    return dict(obs=obs.astype(np.float32), flow=flow.astype(np.float32), depth=depth.astype(np.float32),
                Tcw_init=init.astype(np.float32), Tcw_last=Tlw.astype(np.float32), K=cam, Tcw_gt=Tcw, bad=bad)


def make_pnp(n=1000, seed=0, px_noise=0.1, outliers=0.2, motion_err=0.3, cam=(718.856, 718.856, 607.1928, 185.2157),
             W=1242, H=375):
    """3-D world points of the last frame, their (noisy) projections in the current frame, and a constant-velocity
    guess that is off by `motion_err` metres."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = cam
    Twc = np.eye(4)
    Twc[:3, :3] = ba_synth.rot([0.05, 1, 0.02], 0.08)
    Twc[:3, 3] = [0.4, -0.05, 7.0]
    Tcw = np.linalg.inv(Twc)
    uv = np.stack([rng.uniform(10, W - 10, n), rng.uniform(10, H - 10, n)], 1)
    z = rng.uniform(4, 50, n)
    Xc = np.stack([(uv[:, 0] - cx) * z / fx, (uv[:, 1] - cy) * z / fy, z], 1)
    Xw = Xc @ Twc[:3, :3].T + Twc[:3, 3]
    obs = uv + px_noise * rng.normal(size=(n, 2))
    bad = rng.uniform(size=n) < outliers
    obs[bad] += rng.normal(size=(bad.sum(), 2)) * 15
    mm = Tcw.copy()
    mm[:3, 3] += motion_err * np.array([0.2, 0.05, 1.0])
    return dict(cur=obs.astype(np.float32), pts=Xw.astype(np.float32), Tcw_motion=mm.astype(np.float32), K=cam,
                Tcw_gt=Tcw, bad=bad)
