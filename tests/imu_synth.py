"""Synthetic 200 Hz IMU stream (gravity 9.79, white noise of kaist_config.yaml) and a float64 numpy
integration of the same mid-point scheme used as an independent cross-check."""
import numpy as np

import ba_synth
import oracle_lib as ol

# kaist_config.yaml:105-109 scaled by sqrt(freq) like Tracking::ParseIMUParamFile
FREQ = 200.0
NOISE = (1.7e-4 * np.sqrt(FREQ), 2.0e-3 * np.sqrt(FREQ), 1.9393e-05 / np.sqrt(FREQ), 3.0e-03 / np.sqrt(FREQ))


def make_stream(n_frames=6, seed=0, fps=10.0, jitter=True):
    rng = np.random.default_rng(seed)
    n = int(n_frames / fps * FREQ) + 8
    t = np.arange(n) / FREQ + (0.0013 if jitter else 0.0)
    s = np.zeros(n, ol.IMU_SAMPLE)
    s["t"] = t
    w = np.stack([0.05 * np.sin(2 * t), 0.2 * np.cos(1.3 * t), 0.03 * np.sin(0.7 * t)], 1)
    a = np.stack([0.5 * np.sin(t), -9.79 + 0.1 * np.cos(3 * t), 1.0 + 0.3 * np.sin(2 * t)], 1)
    w += rng.normal(size=w.shape) * 1.7e-4 * np.sqrt(FREQ)
    a += rng.normal(size=a.shape) * 2.0e-3 * np.sqrt(FREQ)
    for k, nm in enumerate(("ax", "ay", "az")):
        s[nm] = a[:, k]
    for k, nm in enumerate(("wx", "wy", "wz")):
        s[nm] = w[:, k]
    frame_t = np.arange(n_frames) / fps + 0.004
    return s, frame_t


def integrate_f64(steps, bias):
    """steps: list of (acc[3], w[3], dt); float64 restatement of dR/dV/dP (no covariance)"""
    dR, dV, dP, dT = np.eye(3), np.zeros(3), np.zeros(3), 0.0
    for a, w, dt in steps:
        acc = np.asarray(a, float) - bias[:3]
        dP = dP + dV * dt + 0.5 * dR @ acc * dt * dt
        dV = dV + dR @ acc * dt
        th = (np.asarray(w, float) - bias[3:]) * dt
        ang = np.linalg.norm(th)
        dRi = ba_synth.rot(th, ang) if ang > 1e-12 else np.eye(3)
        dR = dR @ dRi
        dT += dt
    return dR, dV, dP, dT


# ---------------------------------------------------------------- a consistent body trajectory for the inertial-only optimisation
def _exp(w):
    th = np.linalg.norm(w)
    return ba_synth.rot(w, th) if th > 1e-12 else np.eye(3)


def make_vio_case(n_frames=12, seed=0, fps=10.0, scale_true=1.25, bg_true=(0.002, -0.001, 0.0015), tilt=(0.15, -0.1)):
    """Body trajectory p(t), R(t) in a world whose gravity is tilted by `tilt` (rad about x, y) from -z; 200 Hz IMU samples
    (specific force R^T (a - g), body rates + bias), frames at `fps`.  Returns the inputs of Optimizer::InertialOptimization as
    Tracking::InitializeIMU prepares them (src/Tracking.cc:937-1000): positions in visual units (metric / scale_true), velocities
    from finite differences, Rwg from the summed delta-velocities, zero biases."""
    rng = np.random.default_rng(seed)
    Rg = _exp(np.array([tilt[0], 0, 0])) @ _exp(np.array([0, tilt[1], 0]))
    g_w = Rg @ np.array([0, 0, -9.79])
    h = 1.0 / 2000.0
    T_end = (n_frames - 1) / fps + 0.05
    n_fine = int(T_end / h) + 2
    tf = np.arange(n_fine) * h
    p = np.stack([1.5 * tf + 0.3 * np.sin(1.1 * tf), 0.4 * np.sin(0.8 * tf), 0.2 * np.cos(0.9 * tf) - 0.2], 1)
    v = np.stack([1.5 + 0.33 * np.cos(1.1 * tf), 0.32 * np.cos(0.8 * tf), -0.18 * np.sin(0.9 * tf)], 1)
    a = np.stack([-0.363 * np.sin(1.1 * tf), -0.256 * np.sin(0.8 * tf), -0.162 * np.cos(0.9 * tf)], 1)
    w_body = np.stack([0.10 * np.sin(1.7 * tf), 0.15 * np.cos(1.2 * tf), 0.08 * np.sin(0.9 * tf + 0.3)], 1)
    R = np.zeros((n_fine, 3, 3)); R[0] = _exp(np.array([0.05, -0.02, 0.1]))
    for i in range(1, n_fine):
        R[i] = R[i - 1] @ _exp(0.5 * (w_body[i - 1] + w_body[i]) * h)
    step = int(round(1.0 / FREQ / h))
    idx = np.arange(0, n_fine, step)
    s = np.zeros(len(idx), ol.IMU_SAMPLE)
    s["t"] = tf[idx]
    f_body = np.einsum("nji,nj->ni", R[idx], a[idx] - g_w)            # R^T (a - g)
    f_body += rng.normal(size=f_body.shape) * 2.0e-3 * np.sqrt(FREQ) * 0.1
    w_meas = w_body[idx] + np.asarray(bg_true) + rng.normal(size=f_body.shape) * 1.7e-4 * np.sqrt(FREQ) * 0.1
    for k, nm in enumerate(("ax", "ay", "az")):
        s[nm] = f_body[:, k]
    for k, nm in enumerate(("wx", "wy", "wz")):
        s[nm] = w_meas[:, k]
    fidx = np.array([int(round((k / fps + 0.01) / h)) for k in range(n_frames)])
    ft = tf[fidx]
    Rwb = R[fidx].astype(np.float32)
    twb = (p[fidx] / scale_true).astype(np.float32)
    pre = np.array([ol.imu_preintegrate(s, ft[k], ft[k + 1], np.zeros(6, np.float32), NOISE) for k in range(n_frames - 1)], ol.IMU_PREINT).reshape(-1)
    # Tracking::InitializeIMU: gravity direction from the summed delta-velocities, velocities from finite differences
    dirG = np.zeros(3, np.float32); vel = np.zeros((n_frames, 3), np.float32)
    for k in range(n_frames - 1):
        dirG -= Rwb[k] @ pre[k]["dV"]
        vk = (twb[k + 1] - twb[k]) / pre[k]["dT"]
        vel[k + 1] = vk; vel[k] = vk
    dirG = dirG / np.linalg.norm(dirG)
    gI = np.array([0, 0, -1.0], np.float32)
    vv = np.cross(gI, dirG); nv = np.linalg.norm(vv); ang = np.arccos(np.dot(gI, dirG))
    Rwg = _exp((vv * ang / nv).astype(np.float64))
    truth = dict(scale=scale_true, bg=np.asarray(bg_true), g_dir=g_w / 9.79, vel=v[fidx])
    return dict(Rwb=Rwb, twb=twb, vel=vel, preint=pre, bias_lin=np.zeros((n_frames - 1, 6), np.float32), Rwg=Rwg), truth


# ---------------------------------------------------------------- a camera + IMU sequence for the VIO mode of the tracker
TBC = np.array([[0.9998, -0.0175, 0.0100, 0.05], [0.0174, 0.9998, 0.0080, -0.03], [-0.0101, -0.0078, 0.9999, 0.02], [0, 0, 0, 1]])


def _orthonormal(T):
    u, _, vt = np.linalg.svd(T[:3, :3])
    T = T.copy(); T[:3, :3] = u @ vt
    return T


def vio_camera_pose_np(k):
    """Twc of (real-valued) frame index k: the canyon drive of synth.camera_pose with speed / height / lateral variation so that
    the accelerometer sees the motion (y down: gravity is +y in this world)"""
    yaw = np.radians(2.0) * np.sin(0.15 * k) + np.radians(1.0) * np.sin(0.5 * k)
    pitch = np.radians(0.6) * np.sin(0.31 * k + 1.0)
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    T = np.eye(4)
    T[:3, :3] = Ry @ Rx
    T[:3, 3] = [1.5 * np.sin(0.05 * k) + 0.3 * np.sin(0.4 * k), 0.12 * np.sin(0.5 * k), 1.0 * k + 0.4 * np.sin(0.6 * k)]
    return T


def vio_camera_pose(k):
    import torch
    return torch.from_numpy(vio_camera_pose_np(float(k)))


def vio_camera_pose_10fps_np(k):
    """the same drive for 10 frames/s (bench.py's VIO leg): accelerations of a car (< 2 m/s^2) at 1 m / frame"""
    yaw = np.radians(2.0) * np.sin(0.15 * k)
    pitch = np.radians(0.3) * np.sin(0.11 * k + 1.0)
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    T = np.eye(4)
    T[:3, :3] = Ry @ Rx
    T[:3, 3] = [1.5 * np.sin(0.05 * k) + 0.6 * np.sin(0.1 * k), 0.25 * np.sin(0.13 * k), 1.0 * k + 1.2 * np.sin(0.12 * k)]
    return T


def vio_camera_pose_10fps(k):
    import torch
    return torch.from_numpy(vio_camera_pose_10fps_np(float(k)))


def make_vio_sequence(n_frames, fps=2.5, bg_true=(0.002, -0.001, 0.0015), seed=0, noise=0.05, t0=10.0, pose_np=None):
    """200 Hz IMU samples of the body (Twb = Twc * Tcb) riding on pose_np(t * fps) (default vio_camera_pose_np), frame k at
    t0 + k / fps.  Returns (samples, frame_t, Tbc float32 4x4, truth)"""
    pose_np = pose_np or vio_camera_pose_np
    rng = np.random.default_rng(seed)
    Tbc = _orthonormal(TBC)
    Tcb = np.linalg.inv(Tbc)
    g_w = np.array([0.0, 9.79, 0.0])
    h = 1e-3
    n = int((n_frames - 1) / fps * FREQ) + 12

    def body(t):
        Twc = pose_np(t * fps)
        Twb = Twc @ Tcb
        return Twb[:3, :3], Twb[:3, 3]

    s = np.zeros(n, ol.IMU_SAMPLE)
    ts = (np.arange(n) - 4) / FREQ + 0.0013
    for i, t in enumerate(ts):
        Rm, pm = body(t - h)
        R0, p0 = body(t)
        Rp, pp = body(t + h)
        a = (pp - 2 * p0 + pm) / (h * h)
        dR = Rm.T @ Rp
        w = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / (4 * h)   # log(dR) / (2h), small angle
        f = R0.T @ (a - g_w) + rng.normal(size=3) * 2.0e-3 * np.sqrt(FREQ) * noise
        wm = w + np.asarray(bg_true) + rng.normal(size=3) * 1.7e-4 * np.sqrt(FREQ) * noise
        s["t"][i] = t0 + t
        s["ax"][i], s["ay"][i], s["az"][i] = f
        s["wx"][i], s["wy"][i], s["wz"][i] = wm
    frame_t = t0 + np.arange(n_frames) / fps
    truth = dict(bg=np.asarray(bg_true), g_w=g_w, Tcb=Tcb)
    return s, frame_t, Tbc.astype(np.float32), truth


def imu_chunks(samples, frame_t):
    """per frame k: the samples System::TrackRGBD would receive with it (those up to the frame's timestamp, plus the first later one)"""
    out, lo = [], 0
    for t in frame_t:
        hi = int(np.searchsorted(samples["t"], t, side="right")) + 1
        hi = min(max(hi, lo), len(samples))
        out.append(samples[lo:hi])
        lo = hi
    return out
