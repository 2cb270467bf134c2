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
