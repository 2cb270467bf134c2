"""CPU: the covariance of the IMU preintegration (IMU::Preintegrated::IntegrateNewMeasurement, src/ImuTypes.cc:245-300) pinned by
something that is not a restatement of its recursion: a Monte-Carlo experiment.  The oracle's 9x9 block C[0:9, 0:9] claims to be
the covariance of the error (delta-phi, delta-v, delta-p) of the preintegrated motion when every integration step's gyro / accel
input carries independent white noise of the calibrated (discrete) standard deviation; here 40 000 noisy copies of the same steps
are integrated in float64 and the sample covariance of that error is compared with it.  The bias random-walk block has a closed
form (steps x sigma^2)."""
import numpy as np

import imu_synth
import oracle_lib as ol
from test_imu_oracle import _steps


def _exp_batch(th):
    """rotation matrices of a batch of rotation vectors [M, 3] (Rodrigues)"""
    ang = np.linalg.norm(th, axis=1)
    a = np.where(ang > 1e-12, np.sin(ang) / np.maximum(ang, 1e-300), 1.0)
    b = np.where(ang > 1e-12, (1 - np.cos(ang)) / np.maximum(ang * ang, 1e-300), 0.5)
    K = np.zeros((len(th), 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -th[:, 2], th[:, 1], th[:, 2], -th[:, 0], -th[:, 1], th[:, 0]
    return np.eye(3)[None] + a[:, None, None] * K + b[:, None, None] * (K @ K)


def _log_batch(R):
    c = np.clip((np.trace(R, axis1=1, axis2=2) - 1) / 2, -1, 1)
    ang = np.arccos(c)
    w = np.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], axis=1) / 2
    f = np.where(ang > 1e-9, ang / np.maximum(np.sin(ang), 1e-300), 1.0)
    return w * f[:, None]


def _integrate_batch(steps, bias, noise_g, noise_a):
    """float64 integration of the steps for a batch of noise realisations: noise_g / noise_a [M, n_steps, 3]"""
    M = noise_g.shape[0]
    dR = np.tile(np.eye(3), (M, 1, 1)); dV = np.zeros((M, 3)); dP = np.zeros((M, 3))
    for k, (a, w, dt) in enumerate(steps):
        dt = float(dt)
        acc = np.asarray(a, float)[None] + noise_a[:, k] - bias[:3]
        Ra = np.einsum("mij,mj->mi", dR, acc)
        dP = dP + dV * dt + 0.5 * Ra * dt * dt
        dV = dV + Ra * dt
        dR = dR @ _exp_batch((np.asarray(w, float)[None] + noise_g[:, k] - bias[3:]) * dt)
    return dR, dV, dP


def test_measurement_covariance_matches_monte_carlo():
    s, ft = imu_synth.make_stream(n_frames=3, seed=11)
    bias = np.array([0.02, -0.01, 0.03, 0.001, -0.002, 0.0005], np.float32)
    ng, na, ngw, naw = [float(v) for v in imu_synth.NOISE]
    r = ol.imu_preintegrate(s, ft[0], ft[1], bias, imu_synth.NOISE)
    steps = _steps(s, ft[0], ft[1])
    n = len(steps)
    assert r["n_steps"] == n
    C = r["C"].reshape(15, 15).astype(float)
    rng = np.random.default_rng(1)
    M = 40000
    zero = np.zeros((1, n, 3))
    R0, V0, P0 = _integrate_batch(steps, bias.astype(float), zero, zero)
    R, V, P = _integrate_batch(steps, bias.astype(float), ng * rng.standard_normal((M, n, 3)), na * rng.standard_normal((M, n, 3)))
    err = np.concatenate([_log_batch(np.swapaxes(R0, 1, 2) @ R), V - V0, P - P0], axis=1)   # [M, 9]
    S = err.T @ err / M
    sd = np.sqrt(np.diag(C[:9, :9]))
    assert np.all(sd > 0)
    # diagonal within 4 % (sampling error of a variance from 40 000 draws is 0.7 %; the rest is the first-order model), every
    # entry within 3 % of the product of the two standard deviations
    assert np.abs(np.diag(S) / np.diag(C[:9, :9]) - 1).max() < 0.04, np.diag(S) / np.diag(C[:9, :9])
    assert np.abs((S - C[:9, :9]) / np.outer(sd, sd)).max() < 0.03
    # bias random walk: n steps of independent increments (ImuTypes.cc:291-292)
    assert np.allclose(np.diag(C)[9:12], n * ngw * ngw, rtol=1e-4) and np.allclose(np.diag(C)[12:15], n * naw * naw, rtol=1e-4)
    assert np.abs(C[9:, :9]).max() == 0 and np.abs(C[9:, 9:] - np.diag(np.diag(C)[9:])).max() == 0
