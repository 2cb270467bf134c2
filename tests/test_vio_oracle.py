"""VIO mode of the oracle tracker (sensor = IMU_RGBD): Tracking::PreintegrateIMU per frame, InitializeIMU after the window
optimisation, Map::ApplyScaledRotation, UpdateFrameIMU, ScaleRefinement (src/Tracking.cc:784-1077, 1452-1480)."""
import numpy as np

import imu_synth
import oracle_lib as ol
import synth
import vio_synth


def _rel(Tcw_a, Tcw_b):
    """pose of camera b in the frame of camera a"""
    return Tcw_a.astype(np.float64) @ np.linalg.inv(Tcw_b.astype(np.float64))


def test_imu_initialisation_on_a_tracked_sequence():
    frames, chunks, ft, Tbc, truth = vio_synth.sequence(14)
    tr, poses, states = vio_synth.run_oracle(frames, chunks, ft, Tbc)
    # frames 1..9: too few frames in the map (nMinF = 10); frame 10: 10 map frames spanning 3.6 s
    assert [s[0] for s in states] == [0] * 10 + [1] * 4
    assert states[0][1] == -1 and all(s[1] == 1 for s in states[1:10]) and states[10][1] == 0
    st = tr.imu_state()
    assert st.init_frame == 10 and st.n_reintegrated == 0
    assert abs(st.scale - 1.0) < 0.01                      # the visual map is metric
    assert np.abs(np.array(st.bg[:]) - truth["bg"]).max() < 1e-4
    assert np.abs(np.array(st.ba[:])).max() < 1e-4          # pinned by the 1e9 prior
    Rwg = np.array(st.Rwg[:]).reshape(3, 3)
    g_dir = Rwg @ np.array([0, 0, -1.0])                   # gravity in the visual world (y down)
    assert np.degrees(np.arccos(g_dir @ np.array([0, 1.0, 0]))) < 1.0
    # after Map::ApplyScaledRotation the map's z axis points against gravity: the camera drives along -y'/x' ... the up
    # direction of every map pose is Rgw * (0, -1, 0) = +z
    P = tr.map_poses().reshape(-1, 4, 4)
    up = P[5][:3, :3] @ np.array([0, -1.0, 0])
    assert up[2] > 0.99
    # velocities (frames 1..10) are metric and in the rotated world: about 2.5 m/s forward
    T, vel, bias = tr.imu_frames()
    assert np.all(np.abs(np.linalg.norm(vel[1:11], axis=1) - 2.5) < 0.6)
    assert np.allclose(bias[1:], bias[1])                  # every frame carries the estimated bias
    # tracking goes on in the new world: frame-to-frame motion still matches the ground truth
    for k in range(11, 14):
        rel = _rel(poses[k - 1], poses[k])
        gt = np.linalg.inv(frames[k - 1]["Twc"].numpy()) @ frames[k]["Twc"].numpy()
        assert np.abs(rel[:3, 3] - gt[:3, 3]).max() < 0.05
    tr.close()


def test_reintegration_when_the_gyro_bias_moves():
    frames, chunks, ft, Tbc, truth = vio_synth.sequence(12, bg_true=(0.03, -0.02, 0.015))
    tr, poses, states = vio_synth.run_oracle(frames, chunks, ft, Tbc)
    st = tr.imu_state()
    assert st.initialized == 1 and st.n_reintegrated == 10   # |bg| > 0.01: every map frame's preintegration is redone
    assert np.abs(np.array(st.bg[:]) - truth["bg"]).max() < 2e-4
    tr.close()


def test_apply_scaled_rotation_keeps_the_map_consistent():
    """Map::ApplyScaledRotation on a VO map: p' = s R p, Twc' = [R Rwc | s R twc]"""
    sc = synth.Scene(cam=synth.SMALL, seed=5)
    cfg = ol.track_config(synth.SMALL, rebuild=0, **vio_synth.CFG)
    tr = ol.OracleTracker(cfg)
    for k in range(4):
        f = sc.frame(k)
        tr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
    P0 = tr.map_poses().reshape(-1, 4, 4).astype(np.float64)
    xy, dep, p3, asso = tr.static_features(2)
    th = 0.3
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    tr.apply_scaled_rotation(R, 1.7)
    P1 = tr.map_poses().reshape(-1, 4, 4).astype(np.float64)
    _, _, q3, _ = tr.static_features(2)
    assert np.abs(q3 - 1.7 * p3 @ R.T).max() < 1e-4
    for a, b in zip(P0, P1):
        assert np.abs(b[:3, :3] - R @ a[:3, :3]).max() < 1e-6 and np.abs(b[:3, 3] - 1.7 * R @ a[:3, 3]).max() < 1e-5
    tr.close()


def test_gravity_initialisation_against_a_numpy_restatement():
    """the first half of Tracking::InitializeIMU (src/Tracking.cc:955-988) restated in numpy on the frame states the oracle recorded
    when it initialised: body rotation / position of every Map frame from its camera pose and Tcb (Frame::GetImuRotation /
    GetImuPosition), the gravity direction as the negated sum of the world-frame preintegrated velocity increments, the finite-
    difference velocities, and the rotation that takes (0, 0, -1) onto that direction (IMU::ExpSO3, src/ImuTypes.cc:38-50)"""
    frames, chunks, ft, Tbc, truth = vio_synth.sequence(12)
    tr, poses, states = vio_synth.run_oracle(frames, chunks, ft, Tbc, record=True)
    log = tr.imu_init_log()
    tr.close()
    assert len(log) == 1
    r = log[0]
    n = len(r["has_pre"])
    assert n == 11 and r["has_pre"][1:].all()
    Tcb = np.linalg.inv(r["Tbc"].astype(np.float64))
    Rwb, pwb = [], []
    for T in r["Tcw"].astype(np.float64):
        Rwc = T[:3, :3].T
        Ow = -Rwc @ T[:3, 3]
        Rwb.append(Rwc @ Tcb[:3, :3]); pwb.append(Rwc @ Tcb[:3, 3] + Ow)
    dirG = np.zeros(3)
    vel = np.zeros((n, 3))
    for i in range(1, n):
        if not r["has_pre"][i]:
            continue
        dirG -= Rwb[i - 1] @ r["dV"][i].astype(np.float64)
        v = (pwb[i] - pwb[i - 1]) / float(r["dT"][i])
        vel[i] = v; vel[i - 1] = v
    dirG /= np.linalg.norm(dirG)
    gI = np.array([0.0, 0.0, -1.0])
    v = np.cross(gI, dirG)
    ang = np.arccos(gI @ dirG)
    w = v * ang / np.linalg.norm(v)
    d = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    Rwg = np.eye(3) + W * np.sin(d) / d + W @ W * (1 - np.cos(d)) / (d * d)
    assert np.abs(Rwg - r["Rwg"]).max() < 5e-6
    assert np.abs(Rwg @ gI - dirG).max() < 1e-9                      # it does rotate "down" onto the estimated direction
    assert np.abs(vel - r["vel_out"]).max() < 2e-4 * np.abs(vel).max()
    assert np.degrees(np.arccos((Rwg @ gI) @ np.array([0, 1.0, 0]))) < 1.5   # the synthetic camera's y axis points down


def test_update_frame_imu_against_numpy():
    """Tracking::UpdateFrameIMU (src/Tracking.cc:889-925) on the states the oracle recorded after its initialisation and scale
    refinements: the current frame's body pose and velocity predicted from the previous frame with the bias-corrected
    preintegration, Rwb = Rwb1 dR, twb = twb1 + Vwb1 t + 1/2 t^2 g + Rwb1 dP, Vwb = Vwb1 + g t + Rwb1 dV, g = (0, 0, -9.79)"""
    frames, chunks, ft, Tbc, truth = vio_synth.sequence(14)
    tr, poses, states = vio_synth.run_oracle(frames, chunks, ft, Tbc, record=True)
    log = tr.imu_update_log()
    tr.close()
    assert len(log) >= 1
    g = np.array([0, 0, -9.79])
    for r in log:
        t = float(r["t12"])
        assert 0.05 < t < 1.0
        R1 = r["Rwb1"].astype(np.float64)
        assert np.abs(R1 @ r["dR"].astype(np.float64) - r["Rwb"]).max() < 2e-6
        tw = r["twb1"] + r["Vwb1"].astype(np.float64) * t + 0.5 * t * t * g + R1 @ r["dP"].astype(np.float64)
        vw = r["Vwb1"] + g * t + R1 @ r["dV"].astype(np.float64)
        assert np.abs(tw - r["twb"]).max() < 5e-6 * max(1.0, np.abs(tw).max())
        assert np.abs(vw - r["Vwb"]).max() < 5e-6 * max(1.0, np.abs(vw).max())
