"""CPU: an independent re-solve of the sliding-window graph, the way the reference's linear solver sees it.

The reference hands the FULL system (poses + points, no Schur complement) to CSparse's sparse Cholesky
(3rdparty/g2o/g2o/solvers/linear_solver_csparse.h:108-141).  The oracle (oracle/ba_oracle.cc) and the product eliminate the
points first.  Here one Levenberg-Marquardt trial is restated with numpy / scipy.sparse only: the full H = sum rho' J^T Omega J
and b = -sum rho' J^T Omega e are assembled edge by edge (block_solver.hpp:502-560, base_binary_edge.hpp:55-120, Huber
robust_kernel_impl.cpp:78-91), damped with lambda = tau max|H_jj| (optimization_algorithm_levenberg.cpp:61-120), solved with a
sparse direct solver, applied with the vertices' oplus, and the robust chi2 / gain ratio / next lambda are evaluated -- and
compared with the oracle's first iteration.  Only the per-edge error / Jacobian functions are shared with the oracle (they are
checked against central differences in test_ba_oracle.py)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import ba_synth
import oracle_lib as ol


def quat_from_R(R):  # Eigen::Quaterniond(Matrix3d)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s; s = 0.5 / s
        return np.array([w, (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s])
    i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = np.zeros(4)
    q[1 + i] = 0.5 * s; s = 0.5 / s
    q[0] = (R[k, j] - R[j, k]) * s; q[1 + j] = (R[j, i] + R[i, j]) * s; q[1 + k] = (R[k, i] + R[i, k]) * s
    return q


def R_from_quat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def pose_from_f32(T):  # Converter::toSE3Quat: rotation through a unit quaternion with w >= 0
    T = np.asarray(T, np.float32).reshape(4, 4).astype(np.float64)
    q = quat_from_R(T[:3, :3])
    if q[0] < 0:
        q = -q
    q /= np.linalg.norm(q)
    return np.concatenate([R_from_quat(q).reshape(-1), T[:3, 3]])


def huber(e, delta):
    d2 = delta * delta
    if e <= d2:
        return e, 1.0
    s = np.sqrt(e)
    return 2 * s * delta - d2, delta / s


def robust_chi2(X, Z, pts, op, olp, meas, info_cam, info_3d, d_cam, d_3d):
    chi = 0.0
    for i in range(len(X) - 1):
        e, _, _ = ol.edge_se3(X[i], X[i + 1], Z[i])
        chi += huber(info_cam * e @ e, d_cam)[0]
    for o in range(len(op)):
        e, _, _ = ol.edge_se3_pointxyz(X[op[o]], pts[olp[o]], meas[o])
        chi += huber(info_3d * e @ e, d_3d)[0]
    return chi


@pytest.mark.parametrize("W,P,seed", [(6, 60, 1), (10, 150, 2)])
def test_full_system_sparse_solve_matches_oracle_first_iteration(W, P, seed):
    pr = ba_synth.make_window(W=W, P=P, seed=seed, pose_noise=0.03, obs_noise=0.03)
    info_cam, info_3d = 1.0 / float(np.float32(0.0001)), 1.0 / float(np.float32(16.0))
    d_cam = d_3d = float(np.float32(0.01))
    X = [pose_from_f32(T) for T in pr["poses"]]
    Z = [pose_from_f32(T) for T in pr["rel"]]
    pts = pr["points"].astype(np.float64)
    op, olp, meas = pr["obs_pose"], pr["obs_point"], pr["obs_xyz"].astype(np.float64)
    n = 6 * W + 3 * P
    rows, cols, vals = [], [], []
    b = np.zeros(n)

    def add(i0, j0, B):
        for r in range(B.shape[0]):
            for c in range(B.shape[1]):
                rows.append(i0 + r); cols.append(j0 + c); vals.append(B[r, c])

    for i in range(W - 1):
        e, Ji, Jj = ol.edge_se3(X[i], X[i + 1], Z[i])
        w = huber(info_cam * e @ e, d_cam)[1] * info_cam
        a, c = 6 * i, 6 * (i + 1)
        add(a, a, w * Ji.T @ Ji); add(c, c, w * Jj.T @ Jj); add(a, c, w * Ji.T @ Jj); add(c, a, w * Jj.T @ Ji)
        b[a:a + 6] -= w * Ji.T @ e; b[c:c + 6] -= w * Jj.T @ e
    for o in range(len(op)):
        e, Jp, Jl = ol.edge_se3_pointxyz(X[op[o]], pts[olp[o]], meas[o])
        w = huber(info_3d * e @ e, d_3d)[1] * info_3d
        a, c = 6 * int(op[o]), 6 * W + 3 * int(olp[o])
        add(a, a, w * Jp.T @ Jp); add(c, c, w * Jl.T @ Jl); add(a, c, w * Jp.T @ Jl); add(c, a, w * Jl.T @ Jp)
        b[a:a + 6] -= w * Jp.T @ e; b[c:c + 3] -= w * Jl.T @ e
    H = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc()
    chi0 = robust_chi2(X, Z, pts, op, olp, meas, info_cam, info_3d, d_cam, d_3d)
    lam = 1e-5 * np.abs(H.diagonal()).max()
    x = spla.spsolve((H + lam * sp.identity(n, format="csc")).tocsc(), b)
    Xn = [ol.se3_oplus(X[i], x[6 * i:6 * i + 6]) for i in range(W)]
    ptn = pts + x[6 * W:].reshape(-1, 3)
    chi1 = robust_chi2(Xn, Z, ptn, op, olp, meas, info_cam, info_3d, d_cam, d_3d)
    rho = (chi0 - chi1) / (x @ (lam * x + b) + 1e-3)
    assert rho > 0 and chi1 < chi0            # the first trial is accepted on these graphs
    alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
    lam1 = lam * max(1.0 / 3.0, alpha)
    # ---- the oracle's first LM iteration (Schur complement + dense Cholesky of the reduced system)
    poses, rel, opts, its, st = ol.ba_partial(pr["poses"], pr["rel"], pr["points"], op, olp, pr["obs_xyz"], max_iterations=1)
    c, l, trials = st.records()[0]
    assert its == 1 and trials == 1
    assert abs(c - chi1) <= 1e-7 * chi1        # same step => same robust chi2 after it
    assert abs(l - lam1) <= 1e-6 * lam1        # same gain ratio => same damping for the next iteration
    got = np.stack([pose_from_f32(T) for T in poses])
    want = np.stack(Xn)
    assert np.abs(got - want).max() <= 2e-5 * max(np.abs(want).max(), 1.0)   # outputs are float32 matrices
    assert np.abs(opts - ptn).max() <= 2e-5 * max(np.abs(ptn).max(), 1.0)


def _assemble(X, Z, pts, op, olp, meas, W, P, info_cam, info_3d, d_cam, d_3d):
    n = 6 * W + 3 * P
    rows, cols, vals = [], [], []
    b = np.zeros(n)

    def add(i0, j0, B):
        r, c = np.meshgrid(np.arange(B.shape[0]), np.arange(B.shape[1]), indexing="ij")
        rows.extend((i0 + r).reshape(-1)); cols.extend((j0 + c).reshape(-1)); vals.extend(B.reshape(-1))

    for i in range(W - 1):
        e, Ji, Jj = ol.edge_se3(X[i], X[i + 1], Z[i])
        w = huber(info_cam * e @ e, d_cam)[1] * info_cam
        a, c = 6 * i, 6 * (i + 1)
        add(a, a, w * Ji.T @ Ji); add(c, c, w * Jj.T @ Jj); add(a, c, w * Ji.T @ Jj); add(c, a, w * Jj.T @ Ji)
        b[a:a + 6] -= w * Ji.T @ e; b[c:c + 6] -= w * Jj.T @ e
    for o in range(len(op)):
        e, Jp, Jl = ol.edge_se3_pointxyz(X[op[o]], pts[olp[o]], meas[o])
        w = huber(info_3d * e @ e, d_3d)[1] * info_3d
        a, c = 6 * int(op[o]), 6 * W + 3 * int(olp[o])
        add(a, a, w * Jp.T @ Jp); add(c, c, w * Jl.T @ Jl); add(a, c, w * Jp.T @ Jl); add(c, a, w * Jl.T @ Jp)
        b[a:a + 6] -= w * Jp.T @ e; b[c:c + 3] -= w * Jl.T @ e
    return sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc(), b


@pytest.mark.parametrize("W,P,seed", [(6, 60, 1), (8, 100, 3)])
def test_whole_lm_trajectory_replayed_with_a_sparse_full_system(W, P, seed):
    """Every iteration of the window optimisation, not only the first: the Levenberg-Marquardt driver of g2o
    (optimization_algorithm_levenberg.cpp:61-149: gain ratio, lambda * max(1/3, 1 - (2 rho - 1)^3) / lambda * ni, at most ten
    trials), the reference's two stop patches and SparseOptimizerTerminateAction (gain threshold), replayed in Python on the full
    un-Schur'd system with a sparse direct solver.  The oracle (Schur complement, dense LL^T) must report the same number of
    iterations and trials and the same chi2 / lambda after every iteration."""
    pr = ba_synth.make_window(W=W, P=P, seed=seed, pose_noise=0.03, obs_noise=0.03)
    info_cam, info_3d = 1.0 / float(np.float32(0.0001)), 1.0 / float(np.float32(16.0))
    d_cam = d_3d = float(np.float32(0.01))
    X = [pose_from_f32(T) for T in pr["poses"]]
    Z = [pose_from_f32(T) for T in pr["rel"]]
    pts = pr["points"].astype(np.float64)
    op, olp, meas = pr["obs_pose"], pr["obs_point"], pr["obs_xyz"].astype(np.float64)
    n = 6 * W + 3 * P
    chi = lambda Xs, ps: robust_chi2(Xs, Z, ps, op, olp, meas, info_cam, info_3d, d_cam, d_3d)
    poses, rel, opts, its, st = ol.ba_partial(pr["poses"], pr["rel"], pr["points"], op, olp, pr["obs_xyz"])
    want = st.records()
    bp = ol.BaProblem(); ol.lib().vo_ba_default_params(__import__("ctypes").byref(bp))
    max_it, gain_thr = bp.max_iterations, float(bp.gain_threshold)
    lam, ni, nbad, chi_check, last_chi = -1.0, 2.0, 0, 0.0, 0.0
    got, stop, ok, i = [], False, True, 0
    while i < max_it and not stop and ok:
        cur = ini = chi(X, pts)
        H, b = _assemble(X, Z, pts, op, olp, meas, W, P, info_cam, info_3d, d_cam, d_3d)
        if i == 0:
            lam, ni, nbad = 1e-5 * np.abs(H.diagonal()).max(), 2.0, 0
        rho, q = 0.0, 0
        while True:
            x = spla.spsolve((H + lam * sp.identity(n, format="csc")).tocsc(), b)
            Xn = [ol.se3_oplus(X[k], x[6 * k:6 * k + 6]) for k in range(W)]
            pn = pts + x[6 * W:].reshape(-1, 3)
            tmp = chi(Xn, pn)
            rho = (cur - tmp) / (x @ (lam * x + b) + 1e-3)
            if rho > 0 and np.isfinite(tmp):
                lam *= max(1.0 / 3.0, min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)); ni = 2.0
                cur, X, pts = tmp, Xn, pn
            else:
                lam *= ni; ni *= 2
            q += 1
            if not (rho < 0 and q < 10):
                break
        if q == 10 or rho == 0:
            ok = False
        else:
            nbad = nbad + 1 if (ini - cur) * 1e3 < ini else 0
            ok = nbad < 3
        arc = tmp                       # the errors as the last trial left them (a rejected trial is restored AFTER this is read)
        if chi_check < arc and i > 0:
            ok = False
        chi_check = arc
        got.append((cur, lam, q))
        c = chi(X, pts)
        if i == 0:
            last_chi = c
        else:
            gain = (last_chi - c) / c
            last_chi = c
            if 0 <= gain < gain_thr:
                stop = True
        i += 1
    assert its == len(got) == len(want) and its >= 3
    for (c1, l1, t1), (c2, l2, t2) in zip(got, want):
        assert t1 == t2
        assert abs(c1 - c2) <= 1e-6 * max(c2, 1e-12), (c1, c2)
        assert abs(l1 - l2) <= 1e-4 * l2, (l1, l2)


def _lm_replay(chi_fn, assemble_fn, apply_fn, state, n, max_it, gain_thr):
    """g2o's Levenberg-Marquardt driver + the reference's patches + the terminate action (see the window replay above), generic"""
    lam, ni, nbad, chi_check, last_chi = -1.0, 2.0, 0, 0.0, 0.0
    got, stop, ok, i = [], False, True, 0
    while i < max_it and not stop and ok:
        cur = ini = chi_fn(state)
        H, b = assemble_fn(state)
        if i == 0:
            lam, ni, nbad = 1e-5 * np.abs(H.diagonal()).max(), 2.0, 0
        rho, q = 0.0, 0
        while True:
            x = spla.spsolve((H + lam * sp.identity(n, format="csc")).tocsc(), b)
            trial = apply_fn(state, x)
            tmp = chi_fn(trial)
            rho = (cur - tmp) / (x @ (lam * x + b) + 1e-3)
            if rho > 0 and np.isfinite(tmp):
                lam *= max(1.0 / 3.0, min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)); ni = 2.0
                cur, state = tmp, trial
            else:
                lam *= ni; ni *= 2
            q += 1
            if not (rho < 0 and q < 10):
                break
        if q == 10 or rho == 0:
            ok = False
        else:
            nbad = nbad + 1 if (ini - cur) * 1e3 < ini else 0
            ok = nbad < 3
        if chi_check < tmp and i > 0:
            ok = False
        chi_check = tmp
        got.append((cur, lam, q))
        c = chi_fn(state)
        if i == 0:
            last_chi = c
        else:
            gain = (last_chi - c) / c
            last_chi = c
            if 0 <= gain < gain_thr:
                stop = True
        i += 1
    return got, state


def test_full_batch_lm_trajectory_replayed_with_a_sparse_full_system():
    """Optimizer::FullBatchOptimization (src/Optimizer.cc:1235-2178) on a sub-graph of the golden FullBatch graph (camera poses,
    object motions, static points, whole dynamic tracklets with their LandmarkMotionTernaryEdges, odometry / smoothness edges, the
    prior on the first pose): the LM run replayed in Python on the full sparse system must give the oracle's iterations, trials and
    chi2 / lambda after every iteration."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g2o_golden.npz"))
    g = {k: G["before_" + k] for k in ("se3", "points", "e6_i", "e6_j", "e6_kind", "e6_meas", "obs_se3", "obs_point", "obs_kind", "obs_xyz",
                                       "tern_p1", "tern_p2", "tern_h")}
    n_poses = int(G["n_poses"])
    # ---- sub-graph: every 25th static point, every 25th dynamic tracklet (chains of ternary edges) in full
    NP = len(g["points"])
    nxt = -np.ones(NP, np.int64); has_prev = np.zeros(NP, bool)
    for a, b in zip(g["tern_p1"], g["tern_p2"]):
        nxt[a] = b; has_prev[b] = True
    dyn = np.zeros(NP, bool); dyn[g["tern_p1"]] = True; dyn[g["tern_p2"]] = True
    keep = np.zeros(NP, bool)
    keep[np.nonzero(~dyn)[0][::25]] = True
    heads = [p for p in np.nonzero(dyn)[0] if not has_prev[p]]
    for h in heads[::25]:
        p = h
        while p != -1:
            keep[p] = True; p = nxt[p]
    remap = -np.ones(NP, np.int64); remap[keep] = np.arange(keep.sum())
    ob = keep[g["obs_point"]]; te = keep[g["tern_p1"]] & keep[g["tern_p2"]]
    sub = dict(g, points=g["points"][keep], obs_se3=g["obs_se3"][ob], obs_point=remap[g["obs_point"][ob]].astype(np.int32),
               obs_kind=g["obs_kind"][ob], obs_xyz=g["obs_xyz"][ob], tern_p1=remap[g["tern_p1"][te]].astype(np.int32),
               tern_p2=remap[g["tern_p2"][te]].astype(np.int32), tern_h=g["tern_h"][te])
    P, NS = int(keep.sum()), len(sub["se3"])
    assert 200 <= P <= 800 and te.sum() >= 100
    se3o, ptso, its, st = ol.ba_full(sub, n_poses)
    want = st.records()
    # ---- the same run in Python
    f = lambda v: float(np.float32(v))
    info6 = [1.0 / f(0.0001), 1.0 / f(0.001)]; info3 = [1.0 / f(80.0), 1.0 / f(80.0)]; infoT = 1.0 / f(100.0); infoP = f(100000.0)
    d = f(0.01)
    X0 = [pose_from_f32(T) for T in sub["se3"].reshape(-1, 4, 4)]
    Zp = X0[0].copy()
    Z6 = [pose_from_f32(T) for T in sub["e6_meas"].reshape(-1, 4, 4)]
    meas = sub["obs_xyz"].astype(np.float64)
    n = 6 * NS + 3 * P

    def chi_fn(s):
        X, pts = s
        e, _ = ol.edge_se3_prior(X[0], Zp)
        c = infoP * e @ e
        for k in range(len(sub["e6_i"])):
            e, _, _ = ol.edge_se3(X[sub["e6_i"][k]], X[sub["e6_j"][k]], Z6[k])
            c += huber(info6[sub["e6_kind"][k]] * e @ e, d)[0]
        for o in range(len(meas)):
            e, _, _ = ol.edge_se3_pointxyz(X[sub["obs_se3"][o]], pts[sub["obs_point"][o]], meas[o])
            c += huber(info3[sub["obs_kind"][o]] * e @ e, d)[0]
        for t in range(len(sub["tern_p1"])):
            e, _, _ = ol.edge_landmark_motion(X[sub["tern_h"][t]], pts[sub["tern_p1"][t]], pts[sub["tern_p2"][t]])
            c += huber(infoT * e @ e, d)[0]
        return c

    def assemble_fn(s):
        X, pts = s
        rows, cols, vals = [], [], []
        b = np.zeros(n)

        def add(i0, j0, B):
            r, c = np.meshgrid(np.arange(B.shape[0]), np.arange(B.shape[1]), indexing="ij")
            rows.extend((i0 + r).reshape(-1)); cols.extend((j0 + c).reshape(-1)); vals.extend(B.reshape(-1))

        def binary(a, c, Ja, Jc, e, w):
            add(a, a, w * Ja.T @ Ja); add(c, c, w * Jc.T @ Jc); add(a, c, w * Ja.T @ Jc); add(c, a, w * Jc.T @ Ja)
            b[a:a + Ja.shape[1]] -= w * Ja.T @ e; b[c:c + Jc.shape[1]] -= w * Jc.T @ e

        e, J = ol.edge_se3_prior(X[0], Zp)
        add(0, 0, infoP * J.T @ J); b[0:6] -= infoP * J.T @ e
        for k in range(len(sub["e6_i"])):
            i, j, kd = int(sub["e6_i"][k]), int(sub["e6_j"][k]), int(sub["e6_kind"][k])
            e, Ji, Jj = ol.edge_se3(X[i], X[j], Z6[k])
            binary(6 * i, 6 * j, Ji, Jj, e, huber(info6[kd] * e @ e, d)[1] * info6[kd])
        for o in range(len(meas)):
            v, l, kd = int(sub["obs_se3"][o]), int(sub["obs_point"][o]), int(sub["obs_kind"][o])
            e, Jp, Jl = ol.edge_se3_pointxyz(X[v], pts[l], meas[o])
            binary(6 * v, 6 * NS + 3 * l, Jp, Jl, e, huber(info3[kd] * e @ e, d)[1] * info3[kd])
        for t in range(len(sub["tern_p1"])):
            p1, p2, h = int(sub["tern_p1"][t]), int(sub["tern_p2"][t]), int(sub["tern_h"][t])
            e, J2, JH = ol.edge_landmark_motion(X[h], pts[p1], pts[p2])
            w = huber(infoT * e @ e, d)[1] * infoT
            J1 = np.eye(3)
            blocks = ((6 * NS + 3 * p1, J1), (6 * NS + 3 * p2, J2), (6 * h, JH))
            for oa, Ja in blocks:
                b[oa:oa + Ja.shape[1]] -= w * Ja.T @ e
                for oc, Jc in blocks:
                    add(oa, oc, w * Ja.T @ Jc)
        return sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc(), b

    def apply_fn(s, x):
        X, pts = s
        return [ol.se3_oplus(X[k], x[6 * k:6 * k + 6]) for k in range(NS)], pts + x[6 * NS:].reshape(-1, 3)

    bp = ol.FbaProblem(); ol.lib().vo_fba_default_params(__import__("ctypes").byref(bp))
    got, _ = _lm_replay(chi_fn, assemble_fn, apply_fn, (X0, sub["points"].astype(np.float64)), n, bp.max_iterations, float(bp.gain_threshold))
    assert its == len(got) == len(want) and its >= 3, (its, len(got), len(want))
    for (c1, l1, t1), (c2, l2, t2) in zip(got, want):
        assert t1 == t2
        assert abs(c1 - c2) <= 1e-6 * max(c2, 1e-12), (c1, c2)
        assert abs(l1 - l2) <= 1e-4 * l2, (l1, l2)
