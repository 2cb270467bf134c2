"""CPU: the oracle's restatement of Optimizer::InertialOptimization (src/Optimizer.cc:2441-2620, EdgeInertialGS) on a synthetic
body trajectory with known scale, gravity direction and gyro bias."""
import numpy as np

import imu_synth
import oracle_lib as ol


def test_edge_information_is_the_inverse_of_the_covariance():
    case, _ = imu_synth.make_vio_case(n_frames=4)
    C = case["preint"][1]["C"].reshape(15, 15)[:9, :9].astype(np.float64)
    info = ol.inertial_edge_information(case["preint"][1]["C"])
    assert np.abs(info - info.T).max() <= 1e-9 * np.abs(info).max()
    assert np.all(np.linalg.eigvalsh(info) >= -1e-6 * np.abs(info).max())
    ref = np.linalg.pinv(0.5 * (C + C.T), rcond=2 * 9 * np.finfo(np.float32).eps)
    assert np.abs(info - ref).max() <= 2e-3 * np.abs(ref).max()      # float32 inverse in the reference: 1e-3-level agreement


def test_recovers_scale_gravity_and_gyro_bias():
    case, truth = imu_synth.make_vio_case(n_frames=14, seed=2)
    out = ol.inertial_optimization(**case)
    rec = out["stats"].records()
    assert out["iterations"] >= 3 and rec[-1][0] < 0.05 * rec[0][0]
    assert abs(out["scale"] - truth["scale"]) < 0.03 * truth["scale"], out["scale"]
    g = out["Rwg"] @ np.array([0, 0, -1.0])
    assert np.degrees(np.arccos(np.clip(g @ truth["g_dir"], -1, 1))) < 1.0
    assert np.abs(out["bg"] - truth["bg"]).max() < 6e-4, out["bg"]
    assert np.abs(out["ba"]).max() < 1e-3                              # pinned by the 1e9 prior
    assert np.abs(out["velocity"] * out["scale"] - truth["vel"]).max() < 0.08


def test_scale_refinement_mode():
    """Optimizer::InertialOptimization(Map*, Rwg, scale) of Tracking::ScaleRefinement: only gravity direction and scale move"""
    case, truth = imu_synth.make_vio_case(n_frames=14, seed=2)
    full = ol.inertial_optimization(**case)
    # refine from the initialised state with a deliberately wrong scale, velocities / biases fixed
    out = ol.inertial_optimization(case["Rwb"], case["twb"], full["velocity"], case["preint"], case["bias_lin"], full["Rwg"],
                                   scale=0.9 * full["scale"], bg=full["bg"], ba=full["ba"], mode=1, its=10)
    rec = out["stats"].records()
    assert rec[-1][0] < rec[0][0]
    assert np.array_equal(out["velocity"], full["velocity"]) and np.array_equal(out["bg"], full["bg"])
    assert abs(out["scale"] - full["scale"]) < 0.02 * full["scale"]
