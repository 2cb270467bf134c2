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
