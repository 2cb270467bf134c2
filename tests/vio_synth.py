"""Shared driver of the VIO-mode tests: a 640x480 canyon drive with a 200 Hz IMU riding on the camera (imu_synth.make_vio_sequence)."""
import numpy as np

import imu_synth
import oracle_lib as ol
import synth

CFG = dict(nfeatures=1200, window=8, max_track_bg=400)


def sequence(n_frames, fps=2.5, bg_true=(0.002, -0.001, 0.0015), seed=77, cam=synth.SMALL, **scene_kw):
    samples, ft, Tbc, truth = imu_synth.make_vio_sequence(n_frames, fps=fps, bg_true=bg_true)
    chunks = imu_synth.imu_chunks(samples, ft)
    sc = synth.Scene(cam=cam, seed=seed, pose_fn=imu_synth.vio_camera_pose, **scene_kw)
    frames = [sc.frame(k) for k in range(n_frames)]
    return frames, chunks, ft, Tbc, truth


def run_oracle(frames, chunks, ft, Tbc, cam=synth.SMALL, cfg_kw=None, record=False):
    cfg = ol.track_config(cam, rebuild=0, **(CFG if cfg_kw is None else cfg_kw))
    tr = ol.OracleTracker(cfg)
    if record:
        tr.dyn_log_enable()   # also records the inputs / outputs of the gravity initialisation (OracleTracker.imu_init_log)
    tr.set_imu(Tbc, imu_synth.NOISE)
    poses, states = [], []
    for k, f in enumerate(frames):
        tr.grab_imu(chunks[k])
        T, st, rc = tr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy(), timestamp=ft[k])
        assert rc == 0
        poses.append(T.copy())
        s = tr.imu_state()
        states.append((s.initialized, s.status, s.n_refinements, s.scale))
    return tr, np.stack(poses), states
