"""CPU: the oracle's dynamic-object path (UpdateMask, carry-over, GetSceneFlowObj, DynObjTracking, GetInitModelObj,
PoseOptimizationFlow2, object part of RenewFrameInfo, GetDynamicTrackNew -- src/Tracking.cc:391-421, 1160-1308, 1582-1912,
2030-2162, 2615-2720, 3112-3357) on the synthetic scene with rigid moving billboards, against the generator's ground truth."""
import numpy as np

import oracle_lib as ol
import synth

CAM = synth.KITTI


def _run(n, rebuild=1, **kw):
    sc = synth.Scene(cam=CAM, seed=1234, flow_noise=0.05, depth_noise=0.005, n_objects=5, **kw)
    tr = ol.OracleTracker(ol.track_config(CAM, rebuild=rebuild))
    out = []
    for k in range(n):
        f = sc.frame(k)
        out.append(tr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy()))
    return sc, tr, out


def test_objects_tracked_with_ground_truth_motion():
    n = 6
    sc, tr, out = _run(n)
    T0 = synth.camera_pose(0).numpy()
    ids = {}
    for k in range(1, n):
        st = out[k][1]
        assert st["n_objects"] == 5 and st["n_objects_ok"] == 5, (k, st)
        assert st["n_dyn_features"] >= 150 * 5   # inliers are all kept; 500 per object only gates the top-up
        lab, sem, mot, cen = tr.objects(k)
        assert sorted(sem.tolist()) == [1, 2, 3, 4, 5]
        for l, s, m in zip(lab, sem, mot):
            ids.setdefault(int(s), int(l))
            assert ids[int(s)] == int(l)          # the tracking id of an object never changes
            gt = np.linalg.inv(T0) @ sc.object_motion(int(s) - 1, k - 1).numpy() @ T0   # tracker world = camera-0 frame
            # forward motion of the object (m/frame) to 5 cm, rotation to 1 degree
            assert abs(m[2, 3] - gt[2, 3]) < 0.05, (k, s, m[:3, 3], gt[:3, 3])
            assert np.abs(m[:3, :3] - gt[:3, :3]).max() < 0.02
    assert sorted(ids.values()) == [1, 2, 3, 4, 5]
    ln, oid, ff, fj = tr.dyn_tracks()
    assert len(ln) > 1000 and ln.max() == n and set(np.unique(oid).tolist()) <= {1, 2, 3, 4, 5}
    # every chain starts at a feature that exists
    for t in range(0, len(ln), 97):
        xy = tr.dynamic_features(int(ff[t]))[0]
        assert 0 <= fj[t] < len(xy)
    tr.close()


def test_incremental_tracklets_equal_rebuild():
    _, tr_a, _ = _run(5, rebuild=1)
    _, tr_b, _ = _run(5, rebuild=0)
    for a, b in zip(tr_a.dyn_tracks(), tr_b.dyn_tracks()):
        assert np.array_equal(a, b)
    tr_a.close(); tr_b.close()


def test_lost_mask_is_recovered():
    """the semantic mask of object 2 is missing in frame 3: UpdateMask forward-warps it from frame 2 and the object keeps
    its tracking id"""
    sc, tr, out = _run(5, drop_mask=((3, 2),))
    assert [o[1]["n_masks_recovered"] for o in out] == [0, 0, 0, 1, 0]
    lab3, sem3, _, _ = tr.objects(3)
    lab2, sem2, _, _ = tr.objects(2)
    assert 2 in sem3.tolist()
    assert lab3[sem3.tolist().index(2)] == lab2[sem2.tolist().index(2)]
    tr.close()
