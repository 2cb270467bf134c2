"""GPU: long KITTI-size sequences against the oracle, frame by frame -- the round-off of one frame feeds the decisions of the
next (RANSAC inlier sets, tracklet lengths, which points enter the window), so a long run is the test that small differences
do not compound: 208 static frames, 104 frames with 5 moving objects, and a 1000-frame run whose camera trajectory before
and after FullBatch (the contents of initial_rgbd_new.txt / refined_rgbd_new.txt, src/System.cc:128-198) is compared with
the oracle's.  Frames are fed in chunks (the chunking itself is covered by test_track_gpu.py) so the host never holds more
than one chunk of synthetic input."""
import numpy as np
import pytest

import oracle_lib as ol
import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4
CAM = synth.KITTI
INT_KEYS = ("n_keypoints", "n_matches", "n_init_inliers", "init_winner", "n_pose_inliers", "n_static", "ba_points", "ba_obs")
DYN_KEYS = ("n_dyn_features", "n_objects", "n_objects_ok", "n_masks_recovered")


def _ctx(pkg, batch, **kw):
    return pkg.Context(pkg.default_config(width=CAM["width"], height=CAM["height"], fx=CAM["fx"], fy=CAM["fy"], cx=CAM["cx"],
                                          cy=CAM["cy"], bf=CAM["bf"], max_batch=batch, **kw))


def _run_chunked(pkg, n, chunk, batch, keys, n_objects=0, check_iterations=True):
    """Both trackers over n frames of one scene, chunk frames at a time; every frame's integer statistics must be equal
    and its pose within REL_TOL.  Returns the two trackers (open) for the end-of-run comparisons."""
    dev = "cuda"
    sc = synth.Scene(cam=CAM, seed=4321, flow_noise=0.1, depth_noise=0.01, n_objects=n_objects, device=dev)
    otr = ol.OracleTracker(ol.track_config(CAM))
    ctx = _ctx(pkg, batch)
    worst = 0.0
    for k0 in range(0, n, chunk):
        fr = [sc.frame(k) for k in range(k0, min(k0 + chunk, n))]
        host = [dict(image=f["gray"].cpu().numpy(), depth=f["depth_in"].cpu().numpy(), flow=f["flow"].cpu().numpy(),
                     mask=f["mask"].cpu().numpy()) for f in fr]
        ref = [otr.track(h["image"], h["depth"].copy(), h["flow"], h["mask"].copy()) for h in host]
        T, st = ctx.track_frames([dict(h, mask=h["mask"].copy()) for h in host])
        for i in range(len(fr)):
            k = k0 + i
            T0, s0, rc0 = ref[i]
            assert rc0 == 0, k
            for key in keys:
                assert st[i][key] == s0[key], (k, key, st[i][key], s0[key])
            if check_iterations and s0["ba_points"] >= 50:
                assert (st[i]["ba_iterations"], st[i]["ba_trials"]) == (s0["ba_iterations"], s0["ba_trials"]), k
            err = np.abs(T[i] - T0).max() / max(np.abs(T0).max(), 1.0)
            worst = max(worst, err)
            assert err <= REL_TOL, (k, err)
    return otr, ctx, worst


def _compare_maps(otr, ctx, frames_checked):
    P0, P = otr.map_poses(), ctx.map_poses()
    assert P.shape == P0.shape
    assert np.abs(P - P0).max() <= REL_TOL * max(np.abs(P0).max(), 1.0)
    for fr in frames_checked:
        a, b = ctx.map_static(fr), otr.static_features(fr)
        assert np.array_equal(a[3], b[3]), fr
        assert np.abs(a[2] - b[2]).max() <= REL_TOL * max(np.abs(b[2]).max(), 1.0), fr


def test_static_208_frames_match_oracle(pkg):
    n = 208
    otr, ctx, worst = _run_chunked(pkg, n, 52, 16, INT_KEYS)
    _compare_maps(otr, ctx, (0, 57, n // 2, n - 1))
    print(f"static {n} frames: worst relative pose difference {worst:.2e}")
    otr.close(); ctx.close()


def test_dynamic_104_frames_5_objects_match_oracle(pkg):
    n = 104
    otr, ctx, worst = _run_chunked(pkg, n, 26, 8, INT_KEYS[:3] + INT_KEYS[4:] + DYN_KEYS, n_objects=5)
    _compare_maps(otr, ctx, (0, n // 2, n - 1))
    for k in (1, 40, n - 1):
        a, b = ctx.map_dynamic(k), otr.dynamic_features(k)
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]), k      # vnAssoDyn, vnFeatLabel
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), k
        la, sa, ma, ca = ctx.map_objects(k)
        lb, sb, mb, cb = otr.objects(k)
        assert np.array_equal(la, lb) and np.array_equal(sa, sb), k
        assert np.abs(ma - mb).max() <= REL_TOL * max(np.abs(mb).max(), 1.0), k
    for x, y in zip(ctx.map_dyn_tracks(), otr.dyn_tracks()):
        assert np.array_equal(x, y)
    print(f"dynamic {n} frames: worst relative pose difference {worst:.2e}")
    otr.close(); ctx.close()


def _result_rows(P):
    """The rows SaveResultsIJRR2020 writes per frame (src/System.cc:128-160): frame id, then the 4x4 camera pose row-major."""
    n = P.shape[0]
    return np.concatenate([np.arange(n, dtype=np.float64)[:, None], P.reshape(n, 16).astype(np.float64)], axis=1)


def test_1000_frame_trajectory_files_match_oracle(pkg, tmp_path):
    n = 1000
    otr, ctx, worst = _run_chunked(pkg, n, 50, 16, INT_KEYS)
    ini0, ini = _result_rows(otr.map_poses()), _result_rows(ctx.map_poses())
    otr.full_batch(); ctx.full_batch()
    ref0, ref = _result_rows(otr.map_poses_rf()), _result_rows(ctx.map_poses_rf())
    # the files themselves, the way the facade prints them (one row of 17 numbers per frame)
    for name, a, b in (("initial_rgbd_new.txt", ini0, ini), ("refined_rgbd_new.txt", ref0, ref)):
        np.savetxt(tmp_path / ("oracle_" + name), a, fmt="%.9g")
        np.savetxt(tmp_path / ("b200_" + name), b, fmt="%.9g")
        x, y = np.loadtxt(tmp_path / ("oracle_" + name)), np.loadtxt(tmp_path / ("b200_" + name))
        assert x.shape == y.shape == (n, 17)
        scale = max(np.abs(x[:, 1:]).max(), 1.0)
        diff = np.abs(x - y).max() / scale
        print(f"{name}: {n} rows, largest difference {diff:.2e} of the largest entry ({scale:.1f})")
        assert diff <= REL_TOL
    otr.close(); ctx.close()
