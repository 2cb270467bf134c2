"""GPU: vido_inertial_opt (Optimizer::InertialOptimization, src/Optimizer.cc:2441-2620) against the oracle: same LM trajectory,
scale / gravity direction / biases / velocities within the float tolerance."""
import numpy as np
import pytest

import imu_synth
import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_frames,seed,scale", [(12, 0, 1.25), (20, 3, 0.8), (3, 1, 1.0)])
def test_inertial_opt_matches_oracle(pkg, n_frames, seed, scale):
    case, truth = imu_synth.make_vio_case(n_frames=n_frames, seed=seed, scale_true=scale)
    ref = ol.inertial_optimization(**case)
    ctx = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    out = ctx.inertial_opt(**case)
    a, b = out["stats"], ref["stats"]
    assert a.iterations == b.iterations and a.total_trials == b.total_trials, (a.iterations, b.iterations, a.total_trials, b.total_trials)
    for (c1, l1, t1), (c2, l2, t2) in zip(a.records(), b.records()):
        assert t1 == t2 and abs(c1 - c2) <= 1e-5 * max(abs(c2), 1e-9) and abs(l1 - l2) <= 1e-6 * abs(l2)
    assert abs(out["scale"] - ref["scale"]) <= 1e-6 * abs(ref["scale"])
    assert np.abs(out["Rwg"] - ref["Rwg"]).max() <= 1e-6
    assert np.abs(out["bg"] - ref["bg"]).max() <= 1e-7 and np.abs(out["ba"] - ref["ba"]).max() <= 1e-7
    assert np.abs(out["velocity"] - ref["velocity"]).max() <= 1e-5
    if n_frames >= 12:
        assert abs(out["scale"] - truth["scale"]) < 0.03 * truth["scale"]
    ctx.close()


def test_scale_refinement_mode_matches_oracle(pkg):
    case, truth = imu_synth.make_vio_case(n_frames=14, seed=2)
    full = ol.inertial_optimization(**case)
    args = (case["Rwb"], case["twb"], full["velocity"], case["preint"], case["bias_lin"], full["Rwg"])
    kw = dict(scale=0.9 * full["scale"], bg=full["bg"], ba=full["ba"], mode=1, its=10)
    ref = ol.inertial_optimization(*args, **kw)
    ctx = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    out = ctx.inertial_opt(*args, **kw)
    assert out["stats"].iterations == ref["stats"].iterations and out["stats"].total_trials == ref["stats"].total_trials
    assert abs(out["scale"] - ref["scale"]) <= 1e-6 * abs(ref["scale"]) and np.abs(out["Rwg"] - ref["Rwg"]).max() <= 1e-6
    assert np.array_equal(out["velocity"], ref["velocity"])
    ctx.close()
