"""CPU: IMU preintegration restatement (Tracking::PreintegrateIMU + IMU::Preintegrated) sanity against an independent
float64 numpy integration of the same mid-point scheme."""
import numpy as np

import imu_synth
import oracle_lib as ol


def _steps(s, t0, t1):
    """numpy restatement of the queue selection + interpolation (src/Tracking.cc:806-882)"""
    v = [x for x in s if x["t"] >= t0 - 0.001]
    sel = []
    for x in v:
        sel.append(x)
        if x["t"] >= t1 - 0.001:
            break
    n = len(sel) - 1
    A = lambda x: np.array([x["ax"], x["ay"], x["az"]], np.float32)
    Wv = lambda x: np.array([x["wx"], x["wy"], x["wz"]], np.float32)
    out = []
    for i in range(n):
        a0, a1, w0, w1 = A(sel[i]), A(sel[i + 1]), Wv(sel[i]), Wv(sel[i + 1])
        if i == 0 and i < n - 1:
            tab, tini = np.float32(sel[i + 1]["t"] - sel[i]["t"]), np.float32(sel[i]["t"] - t0)
            out.append(((a0 + a1 - (a1 - a0) * (tini / tab)) * 0.5, (w0 + w1 - (w1 - w0) * (tini / tab)) * 0.5, np.float32(sel[i + 1]["t"] - t0)))
        elif i < n - 1:
            out.append(((a0 + a1) * 0.5, (w0 + w1) * 0.5, np.float32(sel[i + 1]["t"] - sel[i]["t"])))
        elif i > 0:
            tab, tend = np.float32(sel[i + 1]["t"] - sel[i]["t"]), np.float32(sel[i + 1]["t"] - t1)
            out.append(((a0 + a1 - (a1 - a0) * (tend / tab)) * 0.5, (w0 + w1 - (w1 - w0) * (tend / tab)) * 0.5, np.float32(t1 - sel[i]["t"])))
        else:
            out.append((a0, w0, np.float32(t1 - t0)))
    return out


def test_preintegration_matches_float64_integration():
    s, ft = imu_synth.make_stream(n_frames=5, seed=3)
    bias = np.array([0.02, -0.01, 0.03, 0.001, -0.002, 0.0005], np.float32)
    for k in range(1, 5):
        r = ol.imu_preintegrate(s, ft[k - 1], ft[k], bias, imu_synth.NOISE)
        steps = _steps(s, ft[k - 1], ft[k])
        assert r["n_steps"] == len(steps) and 18 <= len(steps) <= 22
        dR, dV, dP, dT = imu_synth.integrate_f64(steps, bias.astype(float))
        assert abs(r["dT"] - dT) < 1e-5 and abs(dT - 0.1) < 1e-5
        assert np.abs(r["dR"].reshape(3, 3) - dR).max() < 2e-6
        assert np.abs(r["dV"] - dV).max() < 2e-5 and np.abs(r["dP"] - dP).max() < 2e-6
        R = r["dR"].reshape(3, 3).astype(float)
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6
        C = r["C"].reshape(15, 15)
        assert np.all(np.diag(C) > 0) and np.abs(C[:9, :9] - C[:9, :9].T).max() < 1e-9 + 1e-5 * np.abs(C).max()


def test_bias_jacobians_predict_rebiased_integration():
    """JRg/JVg/JVa/JPg/JPa are the first-order effect of a bias change (ImuTypes.cc:347-368 uses them that way)"""
    s, ft = imu_synth.make_stream(n_frames=3, seed=5)
    b0 = np.zeros(6, np.float32)
    db = np.array([0.01, -0.02, 0.015, 0.002, -0.001, 0.0015], np.float32)
    r0 = ol.imu_preintegrate(s, ft[0], ft[1], b0, imu_synth.NOISE)
    r1 = ol.imu_preintegrate(s, ft[0], ft[1], b0 + db, imu_synth.NOISE)
    dba, dbg = db[:3].astype(float), db[3:].astype(float)
    dV_pred = r0["dV"] + r0["JVg"].reshape(3, 3) @ dbg + r0["JVa"].reshape(3, 3) @ dba
    dP_pred = r0["dP"] + r0["JPg"].reshape(3, 3) @ dbg + r0["JPa"].reshape(3, 3) @ dba
    assert np.abs(dV_pred - r1["dV"]).max() < 2e-4 and np.abs(dP_pred - r1["dP"]).max() < 2e-5


def test_queue_edge_cases():
    s, ft = imu_synth.make_stream(n_frames=3, seed=1)
    r = ol.imu_preintegrate(s[:1], ft[0], ft[1], np.zeros(6, np.float32), imu_synth.NOISE)
    assert r["n_steps"] == 0 and np.array_equal(r["dR"].reshape(3, 3), np.eye(3, dtype=np.float32))
    two = s[(s["t"] >= ft[0] - 0.001)][:1].copy()
    two = np.concatenate([two, s[s["t"] >= ft[1] - 0.001][:1]])
    r = ol.imu_preintegrate(two, ft[0], ft[1], np.zeros(6, np.float32), imu_synth.NOISE)
    assert r["n_steps"] == 1 and abs(r["dT"] - (ft[1] - ft[0])) < 1e-6
