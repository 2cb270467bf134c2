"""GPU: the per-frame driver (vido_track_frames: batched front-end + sequential back-end incl. the window BA of every
frame) against the oracle's restatement of Tracking::GrabImageRGBD / Track on seeded synthetic sequences."""
import numpy as np
import pytest

import oracle_lib as ol
import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def _sequence(cam, seed, n, **kw):
    sc = synth.Scene(cam=cam, seed=seed, **kw)
    return [sc.frame(k) for k in range(n)]


def _run_both(pkg, cam, frames, max_batch, nfeatures=2500, window=20):
    otr = ol.OracleTracker(ol.track_config(cam, nfeatures=nfeatures, window=window))
    ref = [otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy()) for f in frames]
    ctx = pkg.Context(pkg.default_config(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"],
                                         cy=cam["cy"], bf=cam["bf"], max_batch=max_batch, nfeatures=nfeatures,
                                         window_size=window))
    T, st = ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(),
                                   mask=f["mask"].numpy()) for f in frames])
    return otr, ref, ctx, T, st


@pytest.mark.parametrize("cam_name,n,batch", [("small", 10, 4), ("kitti", 26, 8)])
def test_sequence_matches_oracle(pkg, cam_name, n, batch):
    cam = synth.SMALL if cam_name == "small" else synth.KITTI
    frames = _sequence(cam, 1234, n, flow_noise=0.1, depth_noise=0.01)
    otr, ref, ctx, T, st = _run_both(pkg, cam, frames, batch)
    for k in range(n):
        T0, s0, rc0 = ref[k]
        assert rc0 == 0
        for key in ("n_keypoints", "n_matches", "n_init_inliers", "init_winner", "n_pose_inliers", "n_static",
                    "ba_points", "ba_obs"):
            assert st[k][key] == s0[key], (k, key, st[k][key], s0[key])
        if s0["ba_points"] >= 50:  # tiny odometry-only graphs have chi2 ~ round-off: their stop decisions are noise
            for key in ("ba_iterations", "ba_trials"):
                assert st[k][key] == s0[key], (k, key, st[k][key], s0[key])
        assert np.abs(T[k] - T0).max() <= REL_TOL * max(np.abs(T0).max(), 1.0), k
    P0, P = otr.map_poses(), ctx.map_poses()
    assert P.shape == P0.shape
    assert np.abs(P - P0).max() <= REL_TOL * max(np.abs(P0).max(), 1.0)
    for fr in (0, n // 2, n - 1):
        a, b = ctx.map_static(fr), otr.static_features(fr)
        assert np.array_equal(a[3], b[3])                      # vnAssoSta
        assert np.abs(a[0] - b[0]).max() <= 1e-3               # vpFeatSta (refined flows)
        assert np.abs(a[2] - b[2]).max() <= REL_TOL * max(np.abs(b[2]).max(), 1.0)  # vp3DPointSta after BA
    # the trajectory is also right in absolute terms (ground truth of the generator, frame 0 = origin)
    G0 = frames[0]["Twc"].numpy()
    for k in (n - 1,):
        Tcw_gt = np.linalg.inv(np.linalg.inv(G0) @ frames[k]["Twc"].numpy())
        assert np.abs(T[k][:3, 3] - Tcw_gt[:3, 3]).max() < 0.05 * max(1.0, k * 0.1)
    otr.close(); ctx.close()


def test_chunking_does_not_change_results(pkg):
    cam = synth.SMALL
    frames = _sequence(cam, 77, 7, flow_noise=0.05, depth_noise=0.005)
    outs = []
    for batch in (1, 3, 7):
        ctx = pkg.Context(pkg.default_config(width=640, height=480, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"],
                                             bf=cam["bf"], max_batch=batch))
        T, _ = ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(),
                                      mask=f["mask"].numpy()) for f in frames], want_stats=False)
        outs.append((T.copy(), ctx.map_poses().copy()))
        ctx.close()
    for T, P in outs[1:]:
        assert np.array_equal(T, outs[0][0]) and np.array_equal(P, outs[0][1])


def test_prefetch_hint_does_not_change_results(pkg):
    """vido_track_prefetch only moves the host->device copy of the next call's frames onto the copy stream"""
    cam = synth.SMALL
    frames = _sequence(cam, 78, 8, flow_noise=0.05, depth_noise=0.005)
    host = [dict(image=f["gray"].numpy().copy(), depth=f["depth_in"].numpy().copy(), flow=f["flow"].numpy().copy(),
                 mask=f["mask"].numpy().copy()) for f in frames]
    outs = []
    for hint in (False, True):
        ctx = pkg.Context(pkg.default_config(width=640, height=480, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"],
                                             bf=cam["bf"], max_batch=4))
        if hint:
            ctx.track_prefetch(host[4:8])
        T0, _ = ctx.track_frames(host[0:4], want_stats=False)
        T1, _ = ctx.track_frames(host[4:8], want_stats=False)
        outs.append((np.concatenate([T0, T1]).copy(), ctx.map_poses().copy()))
        ctx.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_rgb_input_and_depth_write_back(pkg):
    """3-channel input goes through the gray conversion kernel; write_back_depth reproduces the reference's in-place
    depth pre-scale"""
    cam = synth.SMALL
    frames = _sequence(cam, 5, 3)
    ctx = pkg.Context(pkg.default_config(width=640, height=480, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"],
                                         bf=cam["bf"], max_batch=2))
    gray = [f["gray"].numpy() for f in frames]
    bgr = [np.repeat(g[:, :, None], 3, axis=2) for g in gray]   # B=G=R=g  ->  gray' == g for the OpenCV coefficients
    deps = [f["depth_in"].numpy().copy() for f in frames]
    T3, _ = ctx.track_frames([dict(image=b, depth=d, flow=f["flow"].numpy(), mask=f["mask"].numpy(), write_back_depth=1)
                              for b, d, f in zip(bgr, deps, frames)], want_stats=False)
    ctx.track_reset()
    T1, _ = ctx.track_frames([dict(image=g, depth=f["depth_in"].numpy(), flow=f["flow"].numpy(), mask=f["mask"].numpy())
                              for g, f in zip(gray, frames)], want_stats=False)
    assert np.array_equal(T1, T3)
    ref = ol.depth_prep(frames[1]["depth_in"].numpy(), 2, 256.0, cam["bf"])
    assert np.array_equal(deps[1], ref)
    ctx.close()


def test_lost_tracking_skips_frames_like_oracle(pkg):
    """a frame whose flow throws every correspondence out of the image leaves mpLastFrame without features: Tracking::Track returns
    early for every later frame (src/Tracking.cc:1110-1113, "Temperal Match size is < 2"), nothing is added to the Map"""
    cam = synth.SMALL
    frames = _sequence(cam, 9, 5)
    flows = [f["flow"].numpy().copy() for f in frames]
    flows[2][:] = 1e4
    otr = ol.OracleTracker(ol.track_config(cam, nfeatures=1200, window=8))
    ref = [otr.track(f["gray"].numpy(), f["depth_in"].numpy(), fl, f["mask"].numpy()) for f, fl in zip(frames, flows)]
    assert [r[2] for r in ref] == [0, 0, 0, 1, 1]
    ctx = pkg.Context(pkg.default_config(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"],
                                         cy=cam["cy"], bf=cam["bf"], max_batch=2, nfeatures=1200, window_size=8))
    T, st = ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=fl, mask=f["mask"].numpy())
                              for f, fl in zip(frames, flows)])
    for k in range(5):
        assert np.abs(T[k] - ref[k][0]).max() <= REL_TOL * max(np.abs(ref[k][0]).max(), 1.0), k
        assert st[k]["n_matches"] == ref[k][1]["n_matches"] and st[k]["n_static"] == ref[k][1]["n_static"], k
    P0, P = otr.map_poses(), ctx.map_poses()
    assert P.shape == P0.shape == (3, 4, 4) and np.abs(P - P0).max() <= REL_TOL
    # the context stays usable: a reset starts a new sequence
    ctx.track_reset()
    T2, _ = ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(), mask=f["mask"].numpy())
                              for f in frames[:3]])
    assert np.abs(T2[1] - ref[1][0]).max() <= REL_TOL
    otr.close(); ctx.close()


def test_cpp_facade_writes_the_reference_result_files(tmp_path, pkg):
    """libvido_slam.so (host/System.cc) frame by frame like run_vido_slam.cc: TrackRGBD x N, then SaveResultsIJRR2020 writes the
    five files of src/System.cc:80-198 in the reference's layout and prints its timing table; the trajectory file equals the
    Map poses the Python binding returns for the same frames; the RGBD overload refuses an IMU_RGBD system (src/System.cc:55)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cam, n = synth.SMALL, 6
    frames = _sequence(cam, 77, n, flow_noise=0.1, depth_noise=0.01)
    raw = tmp_path / "frames.bin"
    with open(raw, "wb") as fh:
        for f in frames:
            fh.write(f["gray"].numpy().tobytes()); fh.write(f["depth_in"].numpy().astype(np.float32).tobytes())
            fh.write(f["flow"].numpy().astype(np.float32).tobytes()); fh.write(f["mask"].numpy().astype(np.int32).tobytes())
    yaml = tmp_path / "cfg.yaml"
    yaml.write_text("%YAML:1.0\n" + "".join(f"Camera.{k}: {cam[k]}\n" for k in ("width", "height", "fx", "fy", "cx", "cy", "bf")) +
                    "ChooseData: 2\nDepthMapFactor: 256.0\n")
    src = tmp_path / "main.cc"
    src.write_text(r'''
#include <cstdio>
#include <cstdlib>
#include "System.h"
int main(int argc, char** argv) {
  const int W = atoi(argv[3]), H = atoi(argv[4]), N = atoi(argv[5]);
  VIDO_SLAM::System sys;
  sys.Init(argv[1], argc > 7 ? VIDO_SLAM::System::IMU_RGBD : VIDO_SLAM::System::RGBD);
  FILE* fh = fopen(argv[2], "rb");
  cv::Mat im = cv::Mat::create(H, W, VIDO_SLAM::CV_8UC1), d = cv::Mat::create(H, W, VIDO_SLAM::CV_32FC1),
          f = cv::Mat::create(H, W, VIDO_SLAM::CV_32FC2), m = cv::Mat::create(H, W, VIDO_SLAM::CV_32SC1), gt, traj;
  std::vector<std::vector<float> > obj;
  for (int k = 0; k < N; k++) {
    if (fread(im.data, 1, (size_t)W * H, fh) != (size_t)W * H) return 2;
    if (fread(d.data, 4, (size_t)W * H, fh) != (size_t)W * H) return 2;
    if (fread(f.data, 8, (size_t)W * H, fh) != (size_t)W * H) return 2;
    if (fread(m.data, 4, (size_t)W * H, fh) != (size_t)W * H) return 2;
    cv::Mat T = sys.TrackRGBD(im, d, f, m, gt, obj, 0.1 * k, traj, N);
    if (T.rows != 4) return 3;
  }
  sys.SaveResultsIJRR2020(argv[6]);
  return 0;
}
''')
    exe = tmp_path / "facade_files"
    libdir = os.path.join(root, "vido-slam_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(libdir, "host"), str(src), "-o", str(exe),
                           "-L", libdir, "-lvido_slam", "-lvido_b200", "-Wl,-rpath," + libdir, "-ldl", "-lpthread", "-lrt"])
    out = tmp_path / "res_"
    args = [str(exe), str(yaml), str(raw), str(cam["width"]), str(cam["height"]), str(n), str(out)]
    run = subprocess.run(args, capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-500:]
    assert "Time of all components:" in run.stdout and "Time of local bundle adjustment:" in run.stdout
    for name in ("obj_mot_rgbd_new.txt", "obj_mot_gt.txt", "initial_rgbd_new.txt", "refined_rgbd_new.txt", "cam_pose_gt.txt"):
        assert os.path.exists(str(out) + name), name
    ini = np.loadtxt(str(out) + "initial_rgbd_new.txt")
    assert ini.shape == (n, 17) and np.array_equal(ini[:, 0], np.arange(n)) and np.all(ini[:, 13:] == [0, 0, 0, 1])
    # the same frames through the Python binding, frame by frame with the facade's settings
    ctx = pkg.Context(pkg.default_config(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"],
                                         bf=cam["bf"], max_batch=1))
    for f in frames:
        ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy().copy(), flow=f["flow"].numpy(), mask=f["mask"].numpy())])
    P = ctx.map_poses().reshape(n, 16)
    assert np.abs(ini[:, 1:13] - P[:, :12]).max() <= 2e-6 * max(np.abs(P).max(), 1.0)   # 9 decimals in the file
    ref = np.loadtxt(str(out) + "refined_rgbd_new.txt")
    assert ref.shape == (n, 17)          # FullBatch ran at the stop frame (ChooseData == 2): refined poses exist for every frame
    gtf = np.loadtxt(str(out) + "cam_pose_gt.txt").reshape(-1, 17)
    assert gtf.shape[0] == 1 and np.array_equal(gtf[0, 1:13], np.eye(4).reshape(16)[:12])
    ctx.close()
    # an IMU_RGBD system must be driven through the IMU overload: the plain overload exits like the reference
    bad = subprocess.run(args + ["imu"], capture_output=True, text=True)
    assert bad.returncode != 0 and "input sensor was not set to RGBD" in bad.stderr


def test_streaming_calls_without_statistics_match_drained_calls(pkg):
    """Calls without a statistics array leave window solves queued (solver host thread + device queue) and the next call
    continues behind them; every observer drains first.  Same Map as one drained call; destroying or resetting a context with
    work still queued neither hangs nor leaks into the next sequence."""
    cam, n = synth.KITTI, 40
    frames = _sequence(cam, 99, n, flow_noise=0.1, depth_noise=0.01)
    host = [dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(), mask=f["mask"].numpy()) for f in frames]
    cfg = dict(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"], bf=cam["bf"], max_batch=8)
    ref = pkg.Context(pkg.default_config(**cfg))
    T0, st0 = ref.track_frames([dict(h, depth=h["depth"].copy()) for h in host])
    P0 = ref.map_poses()
    ref.close()
    ctx = pkg.Context(pkg.default_config(**cfg))
    T = []
    for k0 in range(0, n, 5):   # odd chunking, no statistics: nothing is drained between the calls
        t, st = ctx.track_frames([dict(h, depth=h["depth"].copy()) for h in host[k0:k0 + 5]], want_stats=False)
        assert st is None
        T.append(t)
    assert np.array_equal(np.concatenate(T), T0)
    assert np.array_equal(ctx.map_poses(), P0)            # the accessor retires the queued solves first
    a, b = ctx.map_static(n - 1), None
    ctx.track_reset()                                       # reset right after streaming
    t, _ = ctx.track_frames([dict(h, depth=h["depth"].copy()) for h in host[:12]], want_stats=False)
    assert np.array_equal(t, T0[:12])
    ctx.close()                                             # destroy with solves still queued
    ctx2 = pkg.Context(pkg.default_config(**cfg))
    t2, st2 = ctx2.track_frames([dict(h, depth=h["depth"].copy()) for h in host[:12]])
    assert np.array_equal(t2, T0[:12]) and [s["ba_iterations"] for s in st2] == [s["ba_iterations"] for s in st0[:12]]
    ctx2.close()
