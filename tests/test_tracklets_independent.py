"""CPU: static tracklets and the selection of the window graph, restated independently in pure Python from the reference
(Tracking::GetStaticTrack, src/Tracking.cc:2514-2613; the static part of Optimizer::PartialBatchOptimization's graph
construction, src/Optimizer.cc:46-90, 220-362) and applied to the associations the oracle tracker stored in its Map: for every
frame the number of point vertices and observation edges of the window graph must equal what the oracle (and, by the GPU parity
tests, the product -- which builds tracklets incrementally and selects by 'born inside the window') reports."""
import numpy as np

import oracle_lib as ol
import synth


def get_static_track(temporal_match):
    """temporal_match[i][j]: index in Map frame i of the predecessor of feature j of Map frame i+1 (-1: none)"""
    tracklets, check_pre, ids = [], [], 0
    for i, tm in enumerate(temporal_match):
        check_cur = [-1] * len(tm)
        for j, p in enumerate(tm):
            if p == -1:
                continue
            if i > 0 and check_pre[p] != -1:
                tracklets[check_pre[p]].append((i + 1, j))
                check_cur[j] = check_pre[p]
            else:
                tracklets.append([(i, p), (i + 1, j)])
                check_cur[j] = ids
                ids += 1
        check_pre = check_cur
    return tracklets


def window_graph_size(n_feat, tracklets, window):
    """(point vertices, observation edges) of the window over the last `window` of the len(n_feat) Map frames"""
    N = len(n_feat)
    label = [[-1] * n for n in n_feat]
    for t, tr in enumerate(tracklets):
        if len(tr) < 3:
            continue
        for f, j in tr:
            label[f][j] = t
    mark = [[-1] * n for n in n_feat]
    points = edges = 0
    uid = 0
    for i in range(N - window, N):
        for j in range(n_feat[i]):
            t = label[i][j]
            if t == -1:
                continue
            pos = tracklets[t].index((i, j))
            if pos == 0:
                mark[i][j] = uid; uid += 1
                points += 1; edges += 1
            else:
                pf, pj = tracklets[t][pos - 1]
                if mark[pf][pj] == -1:
                    continue
                mark[i][j] = mark[pf][pj]
                edges += 1
    return points, edges


def test_window_graph_sizes_follow_the_reference_rules():
    cam, n, W = synth.SMALL, 28, 20
    sc = synth.Scene(cam=cam, seed=21, flow_noise=0.1, depth_noise=0.01)
    otr = ol.OracleTracker(ol.track_config(cam, nfeatures=1200, max_track_bg=400, window=W))
    stats = []
    for k in range(n):
        f = sc.frame(k)
        T, s, rc = otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
        assert rc == 0
        stats.append(s)
    feats = [otr.static_features(k) for k in range(n)]
    n_feat = [len(x[3]) for x in feats]
    checked = 0
    for k in range(1, n):
        tm = [feats[f][3].tolist() for f in range(1, k + 1)]          # vnAssoSta as of frame k
        tracklets = get_static_track(tm)
        window = min(k, W)
        pts, obs = window_graph_size(n_feat[:k + 1], tracklets, window)
        assert (stats[k]["ba_points"], stats[k]["ba_obs"]) == (pts, obs), (k, stats[k]["ba_points"], stats[k]["ba_obs"], pts, obs)
        checked += pts > 0
    assert checked >= n - 4 and stats[-1]["ba_points"] > 100
    otr.close()


def test_dynamic_tracklets_follow_the_reference_rules():
    """Tracking::GetDynamicTrackNew (src/Tracking.cc:2615-2720) restated in pure Python on the object associations / labels the
    oracle tracker stored (Map::vnAssoDyn, vnFeatLabel): the same tracklets -- length, object id, first (frame, feature) -- in the
    same order as the oracle's incremental bookkeeping (Map::TrackletDyn, nObjID)"""
    cam, n = synth.SMALL, 12
    sc = synth.Scene(cam=cam, seed=8, flow_noise=0.1, depth_noise=0.01, n_objects=3)
    otr = ol.OracleTracker(ol.track_config(cam, nfeatures=800, max_track_bg=250, max_track_obj=150))
    for k in range(n):
        f = sc.frame(k)
        T, s, rc = otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
        assert rc == 0
    dyn = [otr.dynamic_features(k) for k in range(n)]
    tracklets, obj, check_pre, ids = [], [], [], 0
    for i in range(n - 1):
        asso, lab = dyn[i + 1][3].tolist(), dyn[i + 1][4].tolist()      # vnAssoDyn[i], vnFeatLabel[i]: features of Map frame i + 1
        check_cur = [-1] * len(asso)
        for j, p in enumerate(asso):
            if p == -1:
                continue
            if i > 0 and check_pre[p] != -1:
                tracklets[check_pre[p]].append((i + 1, j))
                check_cur[j] = check_pre[p]
            else:
                tracklets.append([(i, p), (i + 1, j)])
                obj.append(lab[j])
                check_cur[j] = ids
                ids += 1
        check_pre = check_cur
    ln, oid, ff, fj = otr.dyn_tracks()
    assert len(tracklets) == len(ln) and len(tracklets) > 100
    assert [len(t) for t in tracklets] == ln.tolist()
    assert obj == oid.tolist()
    assert [t[0] for t in tracklets] == list(zip(ff.tolist(), fj.tolist()))
    assert max(len(t) for t in tracklets) >= 5
    otr.close()


def test_full_batch_graph_sizes_follow_the_reference_rules():
    """The graph of Optimizer::FullBatchOptimization (src/Optimizer.cc:1235-1760) counted by an independent pure-Python walk over the
    oracle tracker's Map with the reference's rules: camera poses, one motion vertex per tracked object and frame, smoothness edges
    where the label existed in the frame before, static tracklets of length >= 3 (one point, one observation per element), dynamic
    tracklets of length >= 3 (a point and an observation per element, a ternary motion edge per element after the first whose
    object has a motion vertex in that frame).  Must equal the sizes of the graph the oracle exports (and, by the GPU parity
    tests, the product)."""
    cam, n = synth.SMALL, 12
    sc = synth.Scene(cam=cam, seed=8, flow_noise=0.1, depth_noise=0.01, n_objects=3)
    otr = ol.OracleTracker(ol.track_config(cam, nfeatures=800, max_track_bg=250, max_track_obj=150))
    for k in range(n):
        f = sc.frame(k)
        T, s, rc = otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
        assert rc == 0
    g, n_poses = otr.export_full_graph()
    got = (n_poses, len(g["se3"]) - n_poses, len(g["points"]), len(g["obs_se3"]), len(g["e6_i"]), len(g["tern_p1"]))
    # ---- static part
    sta = [otr.static_features(k) for k in range(n)]
    st_tracks = [t for t in get_static_track([sta[f][3].tolist() for f in range(1, n)]) if len(t) >= 3]
    points = len(st_tracks)
    obs = sum(len(t) for t in st_tracks)
    # ---- object motions and smoothness edges
    labels = [None] + [otr.objects(k)[0].tolist() for k in range(1, n)]     # labels[i] = vnRMLabel[i-1][1:], objects of Map frame i
    motions = sum(len(labels[i]) for i in range(1, n))
    smooth = 0
    for i in range(3, n):
        for lab in labels[i]:
            if lab in labels[i - 1]:
                smooth += 1
    e6 = (n - 1) + smooth
    # ---- dynamic part
    dyn = [otr.dynamic_features(k) for k in range(n)]
    ln, oid, ff, fj = otr.dyn_tracks()
    tracklets, check_pre, ids = [], [], 0
    for i in range(n - 1):
        asso = dyn[i + 1][3].tolist()
        check_cur = [-1] * len(asso)
        for j, p in enumerate(asso):
            if p == -1:
                continue
            if i > 0 and check_pre[p] != -1:
                tracklets[check_pre[p]].append((i + 1, j)); check_cur[j] = check_pre[p]
            else:
                tracklets.append([(i, p), (i + 1, j)]); check_cur[j] = ids; ids += 1
        check_pre = check_cur
    tern = 0
    for t, tr in enumerate(tracklets):
        if len(tr) < 3:
            continue
        for pos, (fr, j) in enumerate(tr):
            has_motion = fr >= 1 and int(oid[t]) in labels[fr]
            if pos != 0 and not has_motion:
                continue
            points += 1; obs += 1
            if pos != 0:
                tern += 1
    want = (n, motions, points, obs, e6, tern)
    assert got == want, (got, want)
    assert motions >= 2 * (n - 3) and tern > 100
    # ---- content, independent of the order in which the vertices were numbered
    K = tuple(np.float32(cam[k]) for k in ("fx", "fy", "cx", "cy"))

    def cam_pt(xy, z):   # Optimizer::Get3DinCamera (src/Optimizer.cc:3277-3294)
        return (np.float32(np.float32(np.float32(xy[0] - K[2]) * z) * np.float32(np.float32(1) / K[0])),
                np.float32(np.float32(np.float32(xy[1] - K[3]) * z) * np.float32(np.float32(1) / K[1])), np.float32(z))

    poses = otr.map_poses()
    assert np.array_equal(g["se3"][:n].reshape(n, 4, 4), poses)                       # VertexSE3 estimates = Map::vmCameraPose
    # one vertex per object motion, every one of them started from the identity (src/Optimizer.cc:1577-1583: the estimate of the
    # tracker is commented out there)
    assert np.array_equal(g["se3"][n:].reshape(-1, 16), np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (motions, 1)))
    want_obs, want_pts = [], []
    for t in st_tracks:                                                               # static tracklets: one point, an observation per element
        f0, j0 = t[0]
        want_pts.append(tuple(sta[f0][2][j0]))
        for f, j in t:
            want_obs.append((f,) + cam_pt(sta[f][0][j], sta[f][1][j]))
    want_tern = []
    for t, tr in enumerate(tracklets):                                                # dynamic tracklets: a point and an observation per element
        if len(tr) < 3:
            continue
        for pos, (fr, j) in enumerate(tr):
            has_motion = fr >= 1 and int(oid[t]) in labels[fr]
            if pos != 0 and not has_motion:
                continue
            want_pts.append(tuple(dyn[fr][2][j]))
            want_obs.append((fr,) + cam_pt(dyn[fr][0][j], dyn[fr][1][j]))
            if pos != 0:
                pf, pj = tr[pos - 1]
                want_tern.append(tuple(dyn[pf][2][pj]) + tuple(dyn[fr][2][j]))
    got_obs = sorted((int(a),) + tuple(x) for a, x in zip(g["obs_se3"], g["obs_xyz"].reshape(-1, 3)))
    assert got_obs == sorted(want_obs)
    assert sorted(tuple(x) for x in g["points"].reshape(-1, 3)) == sorted(want_pts)
    P = g["points"].reshape(-1, 3)
    got_tern = sorted(tuple(P[a]) + tuple(P[b]) for a, b in zip(g["tern_p1"], g["tern_p2"]))
    assert got_tern == sorted(want_tern)                                              # ternary edges join point(t-1) and point(t) of a tracklet
    assert (g["tern_h"] >= n).all() and (g["tern_h"] < n + motions).all()             # ... through a motion vertex
    otr.close()


def test_apply_scaled_rotation_against_numpy():
    """Map::ApplyScaledRotation (src/Map.cc:56-119) in numpy on the oracle tracker's Map: points p -> s R p, camera poses
    Twc -> [R | 0] [Rwc | s twc], float32 results"""
    cam, n = synth.SMALL, 5
    sc = synth.Scene(cam=cam, seed=13, flow_noise=0.1, depth_noise=0.01, n_objects=2)
    otr = ol.OracleTracker(ol.track_config(cam, nfeatures=600, max_track_bg=200, max_track_obj=100))
    for k in range(n):
        f = sc.frame(k)
        otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
    P0 = otr.map_poses().astype(np.float64)
    S0 = [otr.static_features(k)[2].astype(np.float64) for k in range(n)]
    D0 = [otr.dynamic_features(k)[2].astype(np.float64) for k in range(n)]
    w = np.array([0.2, -0.1, 0.3]); th = np.linalg.norm(w); K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = (np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K).astype(np.float32)
    s = np.float32(1.37)
    otr.apply_scaled_rotation(R, float(s))
    P1 = otr.map_poses()
    Rd = R.astype(np.float64)
    for k in range(n):
        want = P0[k].copy()
        want[:3, 3] *= float(s)
        want = np.vstack([np.hstack([Rd, np.zeros((3, 1))]), [0, 0, 0, 1]]) @ want
        assert np.abs(P1[k] - want).max() <= 2e-6 * max(np.abs(want).max(), 1.0), k
        a = otr.static_features(k)[2]
        assert np.abs(a - (float(s) * S0[k] @ Rd.T)).max() <= 4e-6 * max(np.abs(S0[k]).max() * float(s), 1.0), k
        b = otr.dynamic_features(k)[2]
        if len(b):
            assert np.abs(b - (float(s) * D0[k] @ Rd.T)).max() <= 4e-6 * max(np.abs(D0[k]).max() * float(s), 1.0), k
    otr.close()
