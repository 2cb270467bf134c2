"""CPU: the oracle's restatement of the reprojection-only optimisers PoseOptimizationNew / PoseOptimizationObjMot
(src/Optimizer.cc:2180-2334, 2826-3035) recovers the true transform and flags the gross outliers."""
import numpy as np

import oracle_lib as ol
import proj_synth


def _run(case):
    d = dict(case)
    kind = d.pop("kind")
    return ol.pose_opt_proj(kind, d["obs_xy"], d["pts3d"], d["T_init"], K=d.get("K"), P=d.get("P"))


def test_camera_pose_recovered():
    case, Tcw, bad = proj_synth.camera_case(seed=1)
    T, inl, st = _run(case)
    assert st.iterations >= 2
    assert np.abs(T - Tcw).max() < 5e-3
    assert not inl[bad].any()                       # every gross outlier is flagged
    assert inl[~bad].mean() > 0.5                   # 0.01 px^2 gate at 0.05 px noise keeps the closest ones


def test_object_motion_recovered():
    case, H, bad = proj_synth.object_case(seed=2, outliers=0.0, noise=0.02)
    T, inl, st = _run(case)
    assert st.iterations >= 2
    assert np.abs(T[:3, 3] - H[:3, 3]).max() < 0.05 and np.abs(T[:3, :3] - H[:3, :3]).max() < 5e-3


def test_degenerate_inputs():
    case, _, _ = proj_synth.camera_case(n=2, seed=3, outliers=0.0)
    T, inl, st = _run(case)
    assert np.array_equal(T, case["T_init"]) and st.iterations == -1
    case, _, _ = proj_synth.object_case(n=2, seed=3, outliers=0.0)
    T, inl, st = _run(case)
    assert np.array_equal(T, np.eye(4, dtype=np.float32))
