"""GPU: single-CTA joint flow+pose optimiser (vido_pose_opt_flow2) against the oracle's restatement of
Optimizer::PoseOptimizationFlow2Cam: same inlier set, same LM trajectory, pose/flows within 1e-4 relative."""
import numpy as np
import pytest

import oracle_lib as ol
import pose_synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    yield c
    c.close()


def _oracle(pr, **params):
    return ol.poseopt_flow2cam(pr["obs"], pr["flow"], pr["depth"], pr["Tcw_init"], pr["Tcw_last"], pr["K"], **params)


def _check(res, ref):
    T, fo, inl, ninl, st = res
    T0, fo0, inl0, ninl0, st0 = ref
    assert ninl == ninl0
    assert np.array_equal(inl, inl0)
    for a, b in zip(st, st0):
        assert a.iterations == b.iterations and a.total_trials == b.total_trials
        for (c1, l1, t1), (c2, l2, t2) in zip(a.records(), b.records()):
            assert t1 == t2 and abs(c1 - c2) <= REL_TOL * max(abs(c2), 1e-9) and abs(l1 - l2) <= 1e-3 * abs(l2)
    assert np.abs(T - T0).max() <= REL_TOL * max(np.abs(T0).max(), 1.0)
    assert np.abs(fo - fo0).max() <= REL_TOL * max(np.abs(fo0).max(), 1.0)


@pytest.mark.parametrize("n,seed,noise,outl", [(900, 1, 0.1, 0.05), (1001, 2, 0.2, 0.15), (60, 3, 0.05, 0.0), (300, 4, 0.5, 0.3)])
def test_flow2cam_matches_oracle(ctx, n, seed, noise, outl):
    pr = pose_synth.make_poseopt(n=n, seed=seed, flow_noise=noise, outliers=outl)
    res = ctx.pose_opt_flow2([dict(obs=pr["obs"], flow=pr["flow"], depth=pr["depth"], Tcw_init=pr["Tcw_init"],
                                   Tcw_last=pr["Tcw_last"], K=pr["K"])])[0]
    _check(res, _oracle(pr))
    assert np.abs(res[0] - pr["Tcw_gt"]).max() < 0.05


def test_batched_problems_and_object_parameters(ctx):
    """several problems in one launch; the object variant uses prior 0.5, one round of 200 iterations"""
    prs = [pose_synth.make_poseopt(n=200 + 50 * k, seed=10 + k) for k in range(5)]
    obj = dict(info_prior=0.5, rounds=1, its=200)
    res = ctx.pose_opt_flow2([dict(obs=p["obs"], flow=p["flow"], depth=p["depth"], Tcw_init=p["Tcw_init"],
                                   Tcw_last=p["Tcw_last"], K=p["K"], params=obj) for p in prs])
    for r, p in zip(res, prs):
        _check(r, _oracle(p, **obj))


def test_fewer_than_three_matches_leaves_everything_untouched(ctx):
    pr = pose_synth.make_poseopt(n=2, seed=5)
    T, fo, inl, ninl, st = ctx.pose_opt_flow2([dict(obs=pr["obs"], flow=pr["flow"], depth=pr["depth"],
                                                    Tcw_init=pr["Tcw_init"], Tcw_last=pr["Tcw_last"], K=pr["K"])])[0]
    assert ninl == 0 and np.array_equal(T, pr["Tcw_init"]) and np.array_equal(fo, pr["flow"])
