"""GPU: the VIO mode of the per-frame driver (vido_track_set_imu / vido_track_grab_imu / vido_track_frames) against the oracle
tracker: preintegration of every frame, InitializeIMU on the inertial kernel, Map::ApplyScaledRotation, UpdateFrameIMU,
ScaleRefinement -- same decisions, poses / velocities / biases within the float tolerances of the stand-alone kernels."""
import numpy as np
import pytest

import imu_synth
import oracle_lib as ol
import synth
import vio_synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def _ctx(pkg, batch):
    cam = synth.SMALL
    ctx = pkg.Context(pkg.default_config(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"],
                                         cy=cam["cy"], bf=cam["bf"], max_batch=batch, nfeatures=vio_synth.CFG["nfeatures"],
                                         window_size=vio_synth.CFG["window"], max_track_bg=vio_synth.CFG["max_track_bg"]))
    return ctx


def _inputs(frames, ft):
    return [dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(), mask=f["mask"].numpy(),
                 timestamp=float(ft[k])) for k, f in enumerate(frames)]


def _compare(ctx, T, otr, poses, n):
    s, r = ctx.imu_state(), otr.imu_state()
    for key in ("initialized", "status", "init_frame", "n_refinements", "n_reintegrated", "lm_iterations", "lm_trials"):
        assert getattr(s, key) == getattr(r, key), (key, getattr(s, key), getattr(r, key))
    assert abs(s.t_init - r.t_init) < 1e-5
    assert abs(s.scale - r.scale) <= 1e-5 * abs(r.scale)
    assert np.abs(np.array(s.Rwg[:]) - np.array(r.Rwg[:])).max() <= 1e-5
    assert np.abs(np.array(s.bg[:]) - np.array(r.bg[:])).max() <= 1e-6 and np.abs(np.array(s.ba[:]) - np.array(r.ba[:])).max() <= 1e-6
    for k in range(n):
        assert np.abs(T[k] - poses[k]).max() <= REL_TOL * max(np.abs(poses[k]).max(), 1.0), k
    Tg, vg, bg_ = ctx.map_imu_frames()
    To, vo, bo = otr.imu_frames()
    assert Tg.shape == To.shape
    assert np.abs(Tg - To).max() <= REL_TOL * max(np.abs(To).max(), 1.0)
    assert np.abs(vg - vo).max() <= 1e-4 * max(np.abs(vo).max(), 1.0)
    assert np.abs(bg_ - bo).max() <= 1e-6
    P0, P = otr.map_poses(), ctx.map_poses()
    assert P.shape == P0.shape and np.abs(P - P0).max() <= REL_TOL * max(np.abs(P0).max(), 1.0)
    a, b = ctx.map_static(n // 2), otr.static_features(n // 2)
    assert np.array_equal(a[3], b[3])
    assert np.abs(a[2] - b[2]).max() <= REL_TOL * max(np.abs(b[2]).max(), 1.0)


@pytest.mark.parametrize("batch,per_call", [(4, 14), (3, 1)])
def test_vio_sequence_matches_oracle(pkg, batch, per_call):
    n = 14
    frames, chunks, ft, Tbc, truth = vio_synth.sequence(n)
    otr, poses, states = vio_synth.run_oracle(frames, chunks, ft, Tbc)
    ctx = _ctx(pkg, batch)
    ctx.track_set_imu(Tbc, imu_synth.NOISE)
    inp = _inputs(frames, ft)
    T = np.zeros((n, 4, 4), np.float32)
    for k0 in range(0, n, per_call):
        k1 = min(n, k0 + per_call)
        T[k0:k1], _ = ctx.track_frames(inp[k0:k1], imu=chunks[k0:k1])
        if per_call == 1:
            s = ctx.imu_state()
            assert (s.initialized, s.status) == states[k0][:2], k0
    assert ctx.imu_state().initialized == 1 and ctx.imu_state().init_frame == 10
    _compare(ctx, T, otr, poses, n)
    otr.close(); ctx.close()


def test_vio_reintegration_matches_oracle(pkg):
    n = 12
    frames, chunks, ft, Tbc, truth = vio_synth.sequence(n, bg_true=(0.03, -0.02, 0.015))
    otr, poses, states = vio_synth.run_oracle(frames, chunks, ft, Tbc)
    ctx = _ctx(pkg, 8)
    ctx.track_set_imu(Tbc, imu_synth.NOISE)
    T, _ = ctx.track_frames(_inputs(frames, ft), imu=chunks)
    assert ctx.imu_state().n_reintegrated == 10
    _compare(ctx, T, otr, poses, n)
    otr.close(); ctx.close()


def test_vio_scale_refinement_matches_oracle(pkg):
    """40 frames at 2.5 fps: mTinit enters (15, 15.5) at frame 38 -> Tracking::ScaleRefinement (mode 1 of the inertial kernel)"""
    n = 40
    frames, chunks, ft, Tbc, truth = vio_synth.sequence(n)
    otr, poses, states = vio_synth.run_oracle(frames, chunks, ft, Tbc)
    assert otr.imu_state().n_refinements == 1
    ctx = _ctx(pkg, 16)
    ctx.track_set_imu(Tbc, imu_synth.NOISE)
    T, _ = ctx.track_frames(_inputs(frames, ft), imu=chunks)
    _compare(ctx, T, otr, poses, n)
    # reset keeps the IMU configuration and starts a new sequence
    ctx.track_reset()
    T2, _ = ctx.track_frames(_inputs(frames[:12], ft[:12]), imu=chunks[:12])
    assert np.array_equal(T2, T[:12])
    otr.close(); ctx.close()


def test_vio_with_dynamic_objects_matches_oracle(pkg):
    """1242x375, 3 moving objects: Map::ApplyScaledRotation also rotates / scales the object points (vp3DPointDyn) and the object
    motions (vmRigidMotion[f][j >= 1]); the object state of mpLastFrame (vObjMod, centres) stays as it is, like the reference"""
    n, cam = 13, synth.KITTI
    frames, chunks, ft, Tbc, truth = vio_synth.sequence(n, cam=cam, seed=1234, flow_noise=0.05, depth_noise=0.005, n_objects=3)
    otr, poses, states = vio_synth.run_oracle(frames, chunks, ft, Tbc, cam=cam, cfg_kw={})
    assert otr.imu_state().initialized == 1
    ctx = pkg.Context(pkg.default_config(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"],
                                         cy=cam["cy"], bf=cam["bf"], max_batch=5))
    ctx.track_set_imu(Tbc, imu_synth.NOISE)
    inp = _inputs(frames, ft)
    for d in inp:
        d["mask"] = d["mask"].copy()
    T, st = ctx.track_frames(inp, imu=chunks)
    # after the initialisation the world's z axis points up: the scene-flow test of DynObjTracking (x / z components only,
    # src/Tracking.cc:1746-1751) no longer sees the objects move -- reference behaviour, identical on both sides
    assert sum(s["n_objects_ok"] for s in st[:11]) >= 3 * 9
    _compare(ctx, T, otr, poses, n)
    for k in (2, 9, 10, n - 1):
        a, b = ctx.map_dynamic(k), otr.dynamic_features(k)
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]), k
        assert np.abs(a[2] - b[2]).max() <= REL_TOL * max(np.abs(b[2]).max(), 1.0), k
        la, sa, ma, ca = ctx.map_objects(k)
        lb, sb, mb, cb = otr.objects(k)
        assert np.array_equal(la, lb) and np.array_equal(sa, sb), k
        if len(mb):   # (none after the initialisation, see above)
            assert np.abs(ma - mb).max() <= REL_TOL * max(np.abs(mb).max(), 1.0), k
    assert len(otr.objects(10)[0]) == 3 and len(otr.objects(n - 1)[0]) == 0
    otr.close(); ctx.close()


def test_apply_scaled_rotation_matches_oracle(pkg):
    sc = synth.Scene(cam=synth.SMALL, seed=5)
    frames = [sc.frame(k) for k in range(4)]
    otr = ol.OracleTracker(ol.track_config(synth.SMALL, rebuild=0, **vio_synth.CFG))
    for f in frames:
        otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
    ctx = _ctx(pkg, 4)
    ctx.track_frames(_inputs(frames, np.arange(4) * 0.1))
    th = 0.3
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    otr.apply_scaled_rotation(R, 1.7); ctx.map_apply_scaled_rotation(R, 1.7)
    P0, P = otr.map_poses(), ctx.map_poses()
    assert np.abs(P - P0).max() <= REL_TOL * max(np.abs(P0).max(), 1.0)
    a, b = ctx.map_static(2), otr.static_features(2)
    assert np.abs(a[2] - b[2]).max() <= REL_TOL * max(np.abs(b[2]).max(), 1.0)
    # the sequence continues from the transformed last frame on both sides
    f = sc.frame(4)
    T0, _, _ = otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
    T, _ = ctx.track_frames(_inputs([f], [0.4]))
    assert np.abs(T[0] - T0).max() <= REL_TOL * max(np.abs(T0).max(), 1.0)
    otr.close(); ctx.close()


def test_cpp_facade_vio_mode(tmp_path, pkg):
    """host/System.h: Init(yaml, IMU_RGBD) parses the Tbc node and the IMU noise like Tracking::ParseIMUParamFile; TrackRGBD with
    vImuMeas feeds Tracking::GrabImuData"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    yaml = tmp_path / "cfg.yaml"
    yaml.write_text("""%YAML:1.0
Camera.width: 640
Camera.height: 480
Camera.fx: 520.0
Camera.fy: 520.0
Camera.cx: 319.5
Camera.cy: 239.5
Camera.bf: 200.0
ChooseData: 2
DepthMapFactor: 256.0
Tbc: !!opencv-matrix
   rows: 4
   cols: 4
   dt: f
   data: [1.0, 0.0, 0.0, 0.05,
          0.0, 1.0, 0.0, -0.03,
          0.0, 0.0, 1.0, 0.02,
          0, 0, 0, 1]

# IMU noise
IMU.NoiseGyro: 1.7e-4
IMU.NoiseAcc: 2.0e-3
IMU.GyroWalk: 1.9393e-05
IMU.AccWalk: 3.0e-03
IMU.Frequency: 200
""")
    src = tmp_path / "main.cc"
    src.write_text(r'''
#include "System.h"
int main(int argc, char** argv) {
  VIDO_SLAM::System sys;
  sys.Init(argv[1], VIDO_SLAM::System::IMU_RGBD);
  VIDO_SLAM::Mat im = VIDO_SLAM::Mat::create(480, 640, VIDO_SLAM::CV_8UC1), d = VIDO_SLAM::Mat::create(480, 640, VIDO_SLAM::CV_32FC1),
                 f = VIDO_SLAM::Mat::create(480, 640, VIDO_SLAM::CV_32FC2), m = VIDO_SLAM::Mat::create(480, 640, VIDO_SLAM::CV_32SC1), gt, traj;
  for (int i = 0; i < 480 * 640; i++) { im.data[i] = (unsigned char)((i * 2654435761u) >> 24); ((float*)d.data)[i] = 256.f * 200.f / 10.f; }
  std::vector<std::vector<float> > obj;
  for (int k = 0; k < 2; k++) {
    std::vector<VIDO_SLAM::IMU::Point> imu;
    for (int j = 0; j < 21; j++) imu.push_back(VIDO_SLAM::IMU::Point(0.f, 9.79f, 0.f, 0.f, 0.f, 0.f, 0.1 * k - 0.1 + 0.005 * j));
    VIDO_SLAM::Mat T = sys.TrackRGBD(im, d, f, m, imu, gt, obj, 0.1 * k, traj, 10);
    if (T.rows != 4) return 2;
  }
  vido_imu_state st;
  if (vido_track_get_imu_state(sys.context(), &st) != VIDO_OK) return 3;   // fails unless Tbc was parsed and IMU mode is on
  float Tcw[32], vel[6], bias[12];
  if (vido_map_get_imu_frames(sys.context(), Tcw, vel, bias, 2) < 1) return 4;
  if (Tcw[3] != 0.f || Tcw[0] != 1.f) return 5;   // Tcb * Tbw of the initial frame = identity for this Tbc
  return sys.isImuInitialized() ? 6 : 0;
}
''')
    exe = tmp_path / "facade_vio"
    libdir = os.path.join(root, "vido-slam_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(libdir, "host"), str(src), "-o", str(exe),
                           "-L", libdir, "-lvido_slam", "-lvido_b200", "-Wl,-rpath," + libdir, "-ldl", "-lpthread", "-lrt"])
    assert subprocess.call([str(exe), str(yaml)]) == 0
