"""CPU: the oracle's restatement of Optimizer::FullBatchOptimization (src/Optimizer.cc:1235-2178) and of the g2o types it adds
to the window graph (LandmarkMotionTernaryEdge, EdgeSE3Prior), on hand-made graphs with ground truth."""
import numpy as np

import fba_synth
import oracle_lib as ol


def _pose12(T):
    return np.concatenate([T[:3, :3].reshape(9), T[:3, 3]])


def test_landmark_motion_edge_matches_its_definition():
    rng = np.random.default_rng(1)
    H = fba_synth._T(fba_synth._rot(0.1, -0.2, 0.3), [0.5, -0.1, 1.2])
    p1, p2 = rng.standard_normal(3), rng.standard_normal(3) + [0, 0, 5]
    e, J2, JH = ol.edge_landmark_motion(_pose12(H), p1, p2)
    q = np.linalg.inv(H)[:3, :3] @ p2 + np.linalg.inv(H)[:3, 3]
    assert np.allclose(e, p1 - q, atol=1e-14)
    assert np.allclose(J2, -H[:3, :3].T, atol=1e-14)          # d err / d p2
    # d err / d (translation increment of VertexSE3::oplus, X <- X * [t, q]) is exactly I ...
    d = 1e-6
    for k in range(3):
        u = np.zeros(6); u[k] = d
        e2 = ol.edge_landmark_motion(ol.se3_oplus(_pose12(H), u), p1, p2)[0]
        assert np.allclose((e2 - e) / d, JH[:, k], atol=1e-6)
    # ... and the rotational columns are the reference's own -[H^-1 p2]x: HALF the derivative w.r.t. the quaternion-vector
    # increment of VertexSE3 (types_dyn_slam3d.cpp:66-76 is restated as written)
    for k in range(3):
        u = np.zeros(6); u[3 + k] = d
        e2 = ol.edge_landmark_motion(ol.se3_oplus(_pose12(H), u), p1, p2)[0]
        assert np.allclose((e2 - e) / d, 2 * JH[:, 3 + k], atol=1e-5)


def test_prior_edge_is_an_se3_edge_from_the_identity():
    X = fba_synth._T(fba_synth._rot(0.2, 0.1, -0.3), [1.0, 2.0, 3.0])
    Z = fba_synth._T(fba_synth._rot(0.21, 0.08, -0.28), [1.1, 1.9, 3.05])
    e, J = ol.edge_se3_prior(_pose12(X), _pose12(Z))
    e2, Ji, Jj = ol.edge_se3(_pose12(np.eye(4)), _pose12(X), _pose12(Z))
    assert np.array_equal(e, e2) and np.array_equal(J, Jj)
    d = 1e-7
    for k in range(6):
        u = np.zeros(6); u[k] = d
        ep = ol.edge_se3_prior(ol.se3_oplus(_pose12(X), u), _pose12(Z))[0]
        assert np.allclose((ep - e) / d, J[:, k], atol=1e-5)


def _pose12_to_T(x):
    T = np.eye(4); T[:3, :3] = np.asarray(x[:9]).reshape(3, 3); T[:3, 3] = x[9:12]
    return T


def test_full_batch_recovers_ground_truth():
    g, n_poses, truth = fba_synth.make_graph(n_frames=6, seed=3)
    se3, pts, its, st = ol.ba_full(g, n_poses)
    rec = st.records()
    assert its >= 3 and all(b[0] <= a[0] * (1 + 1e-12) for a, b in zip(rec, rec[1:]))     # robust chi2 never increases
    before = np.abs(g["se3"].reshape(-1, 4, 4)[:n_poses, :3, 3] - truth["Twc"][:, :3, 3]).max()
    after = np.abs(se3[:n_poses, :3, 3] - truth["Twc"][:, :3, 3]).max()
    assert after < 0.5 * before and after < 0.02
    assert np.array_equal(se3[0], g["se3"].reshape(-1, 4, 4)[0]) or np.abs(se3[0] - g["se3"].reshape(-1, 4, 4)[0]).max() < 1e-5  # prior
    # object motions start from identity and move towards the true forward motion of each object
    for (k, j), vid in truth["motion_vid"].items():
        assert abs(se3[vid][2, 3] - truth["H"][j][k][2, 3]) < 0.15, (k, j, se3[vid][:3, 3], truth["H"][j][k][:3, 3])
    assert np.abs(pts - truth["points"]).max() < 0.2


def test_full_batch_static_only_and_empty():
    g, n_poses, truth = fba_synth.make_graph(n_frames=5, n_objects=0, seed=5)
    se3, pts, its, st = ol.ba_full(g, n_poses)
    assert its >= 2 and np.isfinite(se3).all() and np.isfinite(pts).all()
    assert np.abs(se3[:, :3, 3] - truth["Twc"][:, :3, 3]).max() < 0.02
    empty = {k: np.zeros((0,) + g[k].shape[1:], g[k].dtype) for k in ol.FBA_KEYS}
    assert ol.ba_full(empty, 0)[2] == -1
