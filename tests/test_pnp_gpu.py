"""GPU: initial-model selection (vido_init_model: parallel PnP-RANSAC vs constant-velocity model) against the oracle's
deterministic specification; the RANSAC itself is only loosely comparable with OpenCV (un-vendored, unpinned)."""
import numpy as np
import pytest

import oracle_lib as ol
import pose_synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    yield c
    c.close()


@pytest.mark.parametrize("n,seed,outl,merr", [(1200, 1, 0.2, 0.3), (1200, 2, 0.5, 1.0), (2400, 3, 0.05, 0.0),
                                              (800, 4, 0.7, 0.5), (40, 5, 0.1, 0.2)])
def test_init_model_matches_oracle(ctx, n, seed, outl, merr):
    pr = pose_synth.make_pnp(n=n, seed=seed, outliers=outl, motion_err=merr)
    T0, ids0, w0, nr0, nm0 = ol.init_model_cam(pr["cur"], pr["pts"], None, pr["Tcw_motion"], pr["K"])
    T, ids, w, nr, nm = ctx.init_model(pr["cur"], pr["pts"], None, pr["Tcw_motion"], pr["K"])
    assert w == w0 and nm == nm0
    # FP64 with/without FMA contraction may flip a borderline point at the 0.4 px gate
    assert abs(nr - nr0) <= max(2, 0.005 * nr0)
    common = len(set(ids.tolist()) & set(ids0.tolist()))
    assert common >= 0.995 * max(len(ids0), 1) - 2
    assert np.abs(T - T0).max() <= 1e-4 * max(np.abs(T0).max(), 1.0)
    if w == 0:
        assert np.abs(T - pr["Tcw_gt"]).max() < 0.01


def test_invalid_depths_and_tiny_inputs(ctx):
    pr = pose_synth.make_pnp(n=300, seed=9, outliers=0.1, motion_err=0.4)
    valid = np.ones(300, np.int32)
    valid[::7] = 0
    r0 = ol.init_model_cam(pr["cur"], pr["pts"], valid, pr["Tcw_motion"], pr["K"])
    r1 = ctx.init_model(pr["cur"], pr["pts"], valid, pr["Tcw_motion"], pr["K"])
    assert r1[2] == r0[2] and abs(r1[3] - r0[3]) <= 2 and r1[4] == r0[4]
    pr = pose_synth.make_pnp(n=3, seed=10)
    r0 = ol.init_model_cam(pr["cur"], pr["pts"], None, pr["Tcw_motion"], pr["K"])
    r1 = ctx.init_model(pr["cur"], pr["pts"], None, pr["Tcw_motion"], pr["K"])
    assert r1[2] == r0[2] == 1 and np.array_equal(r1[0], r0[0]) and np.array_equal(r1[1], r0[1])


@pytest.mark.parametrize("k", range(8))
def test_init_model_matches_opencv_golden(ctx, k):
    """the product against the OpenCV goldens (same rule as the CPU test of the oracle), incl. badly wrong motion priors"""
    import test_pnp_golden as tg
    g = np.load(tg.GOLD)
    tg.check_case(g, k, ctx.init_model)
