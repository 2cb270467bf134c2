"""CPU: Tracking::GetMetricError restated (relative camera pose error, body-frame object motion error) on the oracle tracker's Map
against the generator's ground truth."""
import numpy as np

import metric_synth
import oracle_lib as ol
import synth

CAM = synth.KITTI


def test_metric_error_of_a_dynamic_sequence():
    n = 7
    sc = synth.Scene(cam=CAM, seed=1234, flow_noise=0.1, depth_noise=0.01, n_objects=5)
    frames = [sc.frame(k) for k in range(n)]
    tr = ol.OracleTracker(ol.track_config(CAM, rebuild=0))
    for f in frames:
        tr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
    cam_gt, pre, mgt = metric_synth.ground_truth(sc, frames, lambda f: tr.objects(f)[1])
    m, per = tr.metric_error(cam_gt, pre, mgt)
    assert m.n_cam == n - 1 and m.n_obj == 5 * (n - 1)
    assert m.cam_t < 0.02 and m.cam_r < 0.05            # metres / degrees per frame
    assert m.obj_t < 0.03 and m.obj_r < 0.6
    # the averages are the reference's sequential float32 sums of the per-item errors
    ts = np.float32(0)
    for v in per[:n - 1, 0]:
        ts = np.float32(ts + v)
    assert m.cam_t == np.float32(ts / np.float32(n - 1))
    # identical poses give zero error; a known offset comes back as the translation error
    P = tr.map_poses()
    z, _ = tr.metric_error(P)
    assert z.cam_t < 1e-6 and z.cam_r < 0.05 and z.n_obj == 0
    Q = P.copy().reshape(-1, 4, 4)
    Q[3:, 0, 3] += 0.25                                 # ground truth jumps sideways between frames 2 and 3
    _, per2 = tr.metric_error(Q)
    assert abs(per2[2, 0] - 0.25) < 1e-3 and per2[4, 0] < 5e-3   # (a common offset conjugates the later relative poses)
    tr.close()


def test_camera_error_against_a_numpy_restatement():
    """the camera part of Tracking::GetMetricError (src/Tracking.cc:3539-3567) in numpy: ate = (T_i T_{i-1}^-1)(G_{i-1} G_i^-1) on
    float32 matrices, translation norm, rotation angle from the 'folded' trace in degrees with pi = 3.1415926"""
    rng = np.random.default_rng(0)
    sc = synth.Scene(cam=synth.SMALL, seed=3)
    n = 4
    tr = ol.OracleTracker(ol.track_config(synth.SMALL, nfeatures=600, max_track_bg=200))
    for k in range(n):
        f = sc.frame(k)
        tr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
    P = tr.map_poses().reshape(n, 4, 4)                 # Map::vmCameraPose
    G = P.copy()
    for k in range(1, n):                               # a perturbed "ground truth"
        w = rng.normal(0, 0.01, 3); th = np.linalg.norm(w); K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
        G[k, :3, :3] = (R @ G[k, :3, :3].astype(np.float64)).astype(np.float32)
        G[k, :3, 3] += rng.normal(0, 0.05, 3).astype(np.float32)
    _, per = tr.metric_error(G)

    def inv(T):   # Converter::toInvMatrix: [R^T | -R^T t]
        Ti = np.eye(4)
        Ti[:3, :3] = T[:3, :3].T; Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
        return Ti

    for i in range(1, n):
        A = P[i].astype(np.float64) @ inv(P[i - 1].astype(np.float64))
        B = G[i - 1].astype(np.float64) @ inv(G[i].astype(np.float64))
        ate = (A.astype(np.float32).astype(np.float64) @ B.astype(np.float32).astype(np.float64)).astype(np.float32)
        t = np.sqrt(float(ate[0, 3]) ** 2 + float(ate[1, 3]) ** 2 + float(ate[2, 3]) ** 2)
        trace = sum((1.0 - (float(ate[j, j]) - 1.0)) if ate[j, j] > 1.0 else float(ate[j, j]) for j in range(3))
        r = np.degrees(np.arccos(np.clip((trace - 1.0) / 2.0, -1, 1))) * (np.pi / 3.1415926)
        assert abs(per[i - 1, 0] - t) <= 2e-4 * max(t, 1e-3), (i, per[i - 1, 0], t)
        assert abs(per[i - 1, 1] - r) <= 0.06, (i, per[i - 1, 1], r)        # float32 traces near 3: the angle is ill-conditioned
    tr.close()
