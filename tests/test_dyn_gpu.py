"""GPU: vido_track_frames on scenes with dynamic objects against the oracle (same frames, same order): object bookkeeping
is exact (counts, labels, associations, tracklets), object motions / refined features within the float tolerance."""
import numpy as np
import pytest

import metric_synth
import oracle_lib as ol
import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4
CAM = synth.KITTI


def _both(pkg, n, batch, b_joint=1, **kw):
    sc = synth.Scene(cam=CAM, seed=1234, flow_noise=0.1, depth_noise=0.01, n_objects=5, **kw)
    frames = [sc.frame(k) for k in range(n)]
    otr = ol.OracleTracker(ol.track_config(CAM, b_joint=b_joint))
    ref = [otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy()) for f in frames]
    ctx = pkg.Context(pkg.default_config(width=CAM["width"], height=CAM["height"], fx=CAM["fx"], fy=CAM["fy"], cx=CAM["cx"],
                                         cy=CAM["cy"], bf=CAM["bf"], max_batch=batch, b_joint=b_joint))
    T, st = ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(),
                                   mask=f["mask"].numpy().copy()) for f in frames])
    return otr, ref, ctx, T, st


def _compare(otr, ref, ctx, T, st, n):
    for k in range(n):
        T0, s0, rc0 = ref[k]
        assert rc0 == 0
        for key in ("n_keypoints", "n_matches", "n_init_inliers", "n_pose_inliers", "n_static", "ba_points", "ba_obs",
                    "n_dyn_features", "n_objects", "n_objects_ok", "n_masks_recovered"):
            assert st[k][key] == s0[key], (k, key, st[k][key], s0[key])
        assert np.abs(T[k] - T0).max() <= REL_TOL * max(np.abs(T0).max(), 1.0), k
        a, b = ctx.map_dynamic(k), otr.dynamic_features(k)
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]), k      # vnAssoDyn, vnFeatLabel
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), k      # vpFeatDyn (integer pixels / samples), vfDepDyn
        assert np.abs(a[2] - b[2]).max() <= REL_TOL * max(np.abs(b[2]).max(), 1.0), k
        if k >= 1:
            la, sa, ma, ca = ctx.map_objects(k)
            lb, sb, mb, cb = otr.objects(k)
            assert np.array_equal(la, lb) and np.array_equal(sa, sb), k
            assert np.abs(ma - mb).max() <= REL_TOL * max(np.abs(mb).max(), 1.0), k
            assert np.abs(ca - cb).max() <= 1e-4 * max(np.abs(cb).max(), 1.0), k
    for x, y in zip(ctx.map_dyn_tracks(), otr.dyn_tracks()):
        assert np.array_equal(x, y)
    P0, P = otr.map_poses(), ctx.map_poses()
    assert np.abs(P - P0).max() <= REL_TOL * max(np.abs(P0).max(), 1.0)


def test_dynamic_sequence_matches_oracle(pkg):
    n = 12
    otr, ref, ctx, T, st = _both(pkg, n, 4)
    assert all(s["n_objects"] == 5 for s in st[1:])
    _compare(otr, ref, ctx, T, st, n)
    otr.close(); ctx.close()


def test_lost_mask_recovered_like_oracle(pkg):
    n = 6
    otr, ref, ctx, T, st = _both(pkg, n, 3, drop_mask=((3, 2), (4, 5)))
    assert [s["n_masks_recovered"] for s in st] == [0, 0, 0, 1, 1, 0]
    _compare(otr, ref, ctx, T, st, n)
    otr.close(); ctx.close()


def test_reprojection_only_branch_matches_oracle(pkg):
    """bJoint = false (src/Tracking.cc:1133-1136, 1268-1274): PoseOptimizationNew for the camera, PoseOptimizationObjMot per object"""
    n = 8
    otr, ref, ctx, T, st = _both(pkg, n, 4, b_joint=0)
    assert sum(s["n_objects_ok"] for s in st) >= 5 * (n - 2)
    _compare(otr, ref, ctx, T, st, n)
    otr.close(); ctx.close()


def test_metric_error_matches_oracle(pkg):
    """vido_metric_error (Tracking::GetMetricError as a device-side evaluation) on the tracked dynamic sequence, before and after
    the full-sequence optimisation"""
    n = 7
    sc = synth.Scene(cam=CAM, seed=1234, flow_noise=0.1, depth_noise=0.01, n_objects=5)
    frames = [sc.frame(k) for k in range(n)]
    otr = ol.OracleTracker(ol.track_config(CAM, rebuild=0))
    for f in frames:
        otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
    ctx = pkg.Context(pkg.default_config(width=CAM["width"], height=CAM["height"], fx=CAM["fx"], fy=CAM["fy"], cx=CAM["cx"],
                                         cy=CAM["cy"], bf=CAM["bf"], max_batch=4))
    ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(),
                           mask=f["mask"].numpy().copy()) for f in frames])
    cam_gt, pre, mgt = metric_synth.ground_truth(sc, frames, lambda f: otr.objects(f)[1])
    for refined in (False, True):
        if refined:
            otr.full_batch(); ctx.full_batch()
        m0, per0 = otr.metric_error(cam_gt, pre, mgt, refined=refined)
        m1, per1 = ctx.metric_error(cam_gt, pre, mgt, refined=refined)
        assert (m1.n_cam, m1.n_obj) == (m0.n_cam, m0.n_obj) == (n - 1, 5 * (n - 1))
        assert np.abs(per1[:, 0] - per0[:, 0]).max() < 2e-4          # metres
        assert np.abs(per1[:, 1] - per0[:, 1]).max() < 0.06          # degrees (float32 acos near 1: steps of ~0.03 degree)
        assert abs(m1.cam_t - m0.cam_t) < 1e-4 and abs(m1.obj_t - m0.obj_t) < 1e-4
        if not refined:   # (the smoothness edges of the full-sequence optimisation pull the object motions: parity only)
            assert m1.cam_t < 0.02 and m1.obj_t < 0.03
    otr.close(); ctx.close()


@pytest.mark.parametrize("switch", ["VIDO_NO_CHAIN", "VIDO_NO_HYBRID", "VIDO_BA_INLINE"])
def test_paths_agree_under_debug_switches(pkg, monkeypatch, switch):
    """the device-chained tracker, the hybrid object path and the solver host thread are optimisations of the host-driven
    sequential order: switching each of them off gives the same Map (integer bookkeeping equal, floats to 1e-5)"""
    n = 14
    sc = synth.Scene(cam=CAM, seed=777, flow_noise=0.1, depth_noise=0.01, n_objects=3)
    frames = [sc.frame(k) for k in range(n)]

    def run():
        ctx = pkg.Context(pkg.default_config(width=CAM["width"], height=CAM["height"], fx=CAM["fx"], fy=CAM["fy"], cx=CAM["cx"],
                                             cy=CAM["cy"], bf=CAM["bf"], max_batch=4))
        T, st = ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy().copy(), flow=f["flow"].numpy(),
                                       mask=f["mask"].numpy().copy()) for f in frames])
        out = dict(T=T, st=st, P=ctx.map_poses(), dyn=[ctx.map_dynamic(k) for k in (1, n // 2, n - 1)],
                   obj=[ctx.map_objects(k) for k in (1, n // 2, n - 1)], sta=ctx.map_static(n - 1))
        ctx.close()
        return out

    a = run()
    monkeypatch.setenv(switch, "1")
    b = run()
    keys = ("n_keypoints", "n_matches", "n_init_inliers", "n_pose_inliers", "n_static", "ba_points", "ba_obs", "ba_iterations",
            "n_dyn_features", "n_objects", "n_objects_ok", "n_masks_recovered")
    for k in range(n):
        for key in keys:
            assert a["st"][k][key] == b["st"][k][key], (switch, k, key)
    tol = lambda x, y: np.abs(np.asarray(x, np.float64) - np.asarray(y, np.float64)).max() <= 1e-5 * max(np.abs(np.asarray(y)).max(), 1.0)
    assert tol(a["T"], b["T"]) and tol(a["P"], b["P"])
    for da, db in zip(a["dyn"], b["dyn"]):
        assert np.array_equal(da[3], db[3]) and np.array_equal(da[4], db[4]) and np.array_equal(da[0], db[0]) and tol(da[2], db[2])
    for oa, ob in zip(a["obj"], b["obj"]):
        assert np.array_equal(oa[0], ob[0]) and tol(oa[2], ob[2])
    assert np.array_equal(a["sta"][3], b["sta"][3]) and tol(a["sta"][2], b["sta"][2])
