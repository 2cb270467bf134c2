"""CPU: Tracking::RenewFrameInfo (src/Tracking.cc:2959-3289), static half and object half, restated a second time, in Python and from the
reference's text, against the oracle on randomised frame states: inliers carried over through the gates (image bounds with the
reference's strict x <= 0 / y <= 0 test, object mask, depth in (0, 40], both flow components non-zero, correspondence inside the
image), the early exit one PAST the budget, the stride-20 top-up from the detected key points with its 1-pixel "already used" test
against the carried-over set only, truncating depth look-up, back-projection.  The product's feature lists are compared with the
oracle's on the GPU (tests/test_track_gpu.py, tests/test_long_sequence_gpu.py)."""
import numpy as np
import pytest

import oracle_lib as ol

F = np.float32


def renew_static_py(W, H, max_num, K, TM_sta, stat_keys, kps_xy, depth, flow, mask, Tcw):
    keys, corres, fnext, inl = [], [], [], []

    def gates(pt):
        x, y = int(pt[0]), int(pt[1])                      # int x = ...pt.x: truncation
        if x >= W or y >= H or x <= 0 or y <= 0:
            return None
        if mask[y, x] != 0:
            return None
        if depth[y, x] > 40 or depth[y, x] <= 0:
            return None
        fx, fy = flow[y, x]
        if fx != 0 and fy != 0:
            cx, cy = F(pt[0] + fx), F(pt[1] + fy)
            if cx < W and cy < H and cx > 0 and cy > 0:
                return (cx, cy), (fx, fy)
        return None

    for t in TM_sta:                                        # (1) inliers of the last frame
        if t == -1:
            continue
        g = gates(stat_keys[t])
        if g is not None:
            keys.append(tuple(stat_keys[t])); corres.append(g[0]); fnext.append(g[1]); inl.append(int(t))
        if len(keys) > max_num:
            break
    check = np.array(keys, np.float32).reshape(-1, 2)       # mvKeysTmpCheck: a copy, the top-up never tests against itself
    tot, start, step = len(keys), 0, 20
    while tot < max_num:                                    # (2) top-up
        if start == step:
            break
        for i in range(start, len(kps_xy), step):
            s = kps_xy[i]
            if len(check):
                d = np.sqrt((check[:, 0] - s[0]) ** 2 + (check[:, 1] - s[1]) ** 2, dtype=np.float32)
                if (d < 1.0).any():
                    continue
            g = gates(s)
            if g is not None:
                keys.append(tuple(s)); corres.append(g[0]); fnext.append(g[1]); inl.append(-1)
                tot += 1
            if tot >= max_num:
                break
        start += 1
    n = len(keys)
    keys = np.array(keys, np.float32).reshape(n, 2)
    dep = np.full(n, -1, np.float32)                        # (3) depth at the truncated position
    for i in range(n):
        d = depth[int(keys[i, 1]), int(keys[i, 0])]
        if d > 0:
            dep[i] = d
    Twc = np.linalg.inv(Tcw.astype(np.float64))             # (4) Get3DinWorld
    fx, fy, cx, cy = K
    cam = np.stack([(keys[:, 0] - cx) * dep / fx, (keys[:, 1] - cy) * dep / fy, dep], 1).astype(np.float64)
    p3 = cam @ Twc[:3, :3].T + Twc[:3, 3]
    return keys, np.array(corres, np.float32).reshape(n, 2), np.array(fnext, np.float32).reshape(n, 2), np.array(inl, np.int32), dep, p3


@pytest.mark.parametrize("seed,max_num,n_stat,n_kps", [(0, 60, 90, 400), (1, 60, 20, 400), (2, 200, 150, 300), (3, 40, 200, 50),
                                                       (4, 500, 100, 900), (5, 30, 0, 700), (6, 80, 81, 0), (7, 100, 300, 2500)])
def test_static_renewal_matches_an_independent_restatement(seed, max_num, n_stat, n_kps):
    rng = np.random.default_rng(seed)
    W, H = 160, 120
    cam = dict(width=W, height=H, fx=100.0, fy=110.0, cx=80.5, cy=59.5, bf=50.0)
    cfg = ol.track_config(cam, max_track_bg=max_num)
    depth = rng.uniform(1, 50, (H, W)).astype(np.float32)          # a fifth beyond the 40 m gate
    depth[rng.random((H, W)) < 0.05] = 0
    depth[rng.random((H, W)) < 0.02] = -1
    flow = rng.normal(0, 6, (H, W, 2)).astype(np.float32)
    flow[rng.random((H, W)) < 0.05, 0] = 0                         # one zero component disables the feature
    flow[rng.random((H, W)) < 0.05, 1] = 0
    mask = (rng.random((H, W)) < 0.15).astype(np.int32) * rng.integers(1, 4, (H, W)).astype(np.int32)
    stat = np.stack([rng.uniform(-3, W + 3, n_stat), rng.uniform(-3, H + 3, n_stat)], 1).astype(np.float32)
    if n_stat:
        stat[: n_stat // 8] = np.floor(stat[: n_stat // 8])         # integer positions, some exactly on column / row 0
        stat[0] = [0.4, 10.0]
    TM = rng.permutation(n_stat).astype(np.int32)
    TM[rng.random(n_stat) < 0.3] = -1
    kps = np.zeros(n_kps, ol.KP_DTYPE)
    kps["x"] = np.floor(rng.uniform(0, W, n_kps)); kps["y"] = np.floor(rng.uniform(0, H, n_kps))
    if n_stat and n_kps:                                            # detected key points that sit on carried-over features
        k = min(n_kps, n_stat) // 2
        kps["x"][:k] = stat[:k, 0] + rng.uniform(-0.6, 0.6, k).astype(np.float32)
        kps["y"][:k] = stat[:k, 1] + rng.uniform(-0.6, 0.6, k).astype(np.float32)
    a = rng.normal(0, 0.1, 3)
    th = np.linalg.norm(a); k = a / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    T = np.eye(4); T[:3, :3] = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx; T[:3, 3] = rng.normal(0, 1, 3)
    T = T.astype(np.float32)
    want = ol.renew_static(cfg, TM, stat, kps, depth, flow, mask, T)
    got = renew_static_py(W, H, max_num, (F(cam["fx"]), F(cam["fy"]), F(cam["cx"]), F(cam["cy"])), TM, stat,
                          np.stack([kps["x"], kps["y"]], 1), depth, flow, mask, T)
    assert len(got[0]) == len(want[0])
    for g, w, name in zip(got[:5], want[:5], ("keys", "corres", "flow", "inlier id", "depth")):
        assert np.array_equal(g, w), name
    assert np.abs(got[5] - want[5]).max() <= 2e-5 * max(1.0, np.abs(want[5]).max())
    if n_stat >= 3 * max_num:
        assert len(want[0]) == max_num + 1                          # the carry-over loop leaves one past the budget (> not >=)
    if n_stat == 0 and n_kps >= 700:
        assert len(want[0]) == max_num and (want[3] == -1).all()    # filled entirely from the detected key points


def renew_objects_py(W, H, max_num_obj, K, obj_keys, obj_label, inlier_sets, obj_stat, sem_pos, mod_label, tmp, depth, flow, mask, Tcw):
    """object half, src/Tracking.cc:3112-3289"""
    keys, dep, cor, fl, sem, inl, lab = [], [], [], [], [], [], []
    count = []
    for o, ids in enumerate(inlier_sets):                   # (1) inliers of each successfully tracked object, at TRUNCATED positions
        if not obj_stat[o]:
            count.append(-1)
            continue
        c = 0
        for i in ids:
            x, y = int(obj_keys[i][0]), int(obj_keys[i][1])
            if x >= W or y >= H or x <= 0 or y <= 0:
                continue
            if mask[y, x] != 0 and depth[y, x] < 25 and depth[y, x] > 0:
                fx, fy = flow[y, x]
                cx, cy = F(F(x) + fx), F(F(y) + fy)
                if cx < W and cy < H and cx > 0 and cy > 0:
                    keys.append((F(x), F(y))); dep.append(depth[y, x]); sem.append(int(mask[y, x])); fl.append((fx, fy))
                    cor.append((cx, cy)); inl.append(int(i)); lab.append(int(obj_label[i]))
                    c += 1
        count.append(c)
    check = np.array(keys, np.float32).reshape(-1, 2)       # the copy the "already used" test looks at: inliers only
    for o in range(len(inlier_sets)):                       # (2) top-up per object from the frame's samples, 15 interleaved passes
        if not obj_stat[o]:
            continue
        tot, start, step = count[o], 0, 15
        while tot < max_num_obj:
            if start == step:
                break
            for j in range(start, len(tmp["sem"]), step):
                if tmp["sem"][j] != sem_pos[o]:
                    continue
                s = tmp["keys"][j]
                if len(check):
                    d = np.sqrt((check[:, 0] - s[0]) ** 2 + (check[:, 1] - s[1]) ** 2, dtype=np.float32)
                    if (d < 1.0).any():
                        continue
                keys.append(tuple(s)); dep.append(tmp["depth"][j]); sem.append(int(tmp["sem"][j])); fl.append(tuple(tmp["flow"][j]))
                cor.append(tuple(tmp["corres"][j])); inl.append(-1); lab.append(int(mod_label[o]))
                tot += 1
                if tot >= max_num_obj:
                    break
            start += 1
    tracked = set(int(sem_pos[o]) for o in range(len(sem_pos)) if obj_stat[o])
    for u in sorted(set(int(v) for v in tmp["sem"])):       # (3) labels without a tracked object: every sample, label -2
        if u in tracked:
            continue
        for j in range(len(tmp["sem"])):
            if tmp["sem"][j] == u:
                keys.append(tuple(tmp["keys"][j])); dep.append(tmp["depth"][j]); sem.append(u); fl.append(tuple(tmp["flow"][j]))
                cor.append(tuple(tmp["corres"][j])); inl.append(-1); lab.append(-2)
    n = len(keys)
    keys = np.array(keys, np.float32).reshape(n, 2); dep = np.array(dep, np.float32)
    Twc = np.linalg.inv(Tcw.astype(np.float64))
    fx, fy, cx, cy = K
    cam = np.stack([(keys[:, 0] - cx) * dep / fx, (keys[:, 1] - cy) * dep / fy, dep], 1).astype(np.float64)
    p3 = cam @ Twc[:3, :3].T + Twc[:3, 3]
    return (keys, dep, np.array(cor, np.float32).reshape(n, 2), np.array(fl, np.float32).reshape(n, 2), np.array(sem, np.int32),
            np.array(inl, np.int32), np.array(lab, np.int32), p3)


@pytest.mark.parametrize("seed,max_num_obj", [(0, 60), (1, 25), (2, 200), (3, 10), (4, 120), (5, 40)])
def test_object_renewal_matches_an_independent_restatement(seed, max_num_obj):
    rng = np.random.default_rng(100 + seed)
    W, H = 200, 140
    cam = dict(width=W, height=H, fx=100.0, fy=110.0, cx=99.5, cy=69.5, bf=50.0)
    cfg = ol.track_config(cam, max_track_obj=max_num_obj)
    nlab = 5                                                 # semantic labels 1..5 as vertical stripes, label 0 in between
    mask = np.zeros((H, W), np.int32)
    for l in range(nlab):
        mask[20:120, 10 + 38 * l: 10 + 38 * l + 30] = l + 1
    mask[rng.random((H, W)) < 0.03] = 0
    depth = rng.uniform(2, 30, (H, W)).astype(np.float32)    # a fifth beyond the 25 m object gate
    depth[rng.random((H, W)) < 0.03] = 0
    flow = rng.normal(0, 5, (H, W, 2)).astype(np.float32)
    # the frame's object samples (stride-4 grid inside the masks, like Frame::Frame)
    ys, xs = np.mgrid[0:H:4, 0:W:4]
    sel = mask[ys, xs] != 0
    tk = np.stack([xs[sel], ys[sel]], 1).astype(np.float32)
    tmp = dict(keys=tk, depth=depth[ys[sel], xs[sel]], sem=mask[ys[sel], xs[sel]], flow=flow[ys[sel], xs[sel]],
               corres=tk + flow[ys[sel], xs[sel]])
    # the tracked objects of this frame: labels 1..4 (label 5 is "new"), one of them failed; features carried from the last frame
    n_obj = 4
    obj_stat = np.ones(n_obj, np.int32); obj_stat[rng.integers(0, n_obj)] = 0
    sem_pos = np.array([1, 2, 3, 4], np.int32)[rng.permutation(4)]
    mod_label = (rng.permutation(9)[:n_obj] + 1).astype(np.int32)
    n_feat = 500
    ok = np.stack([rng.uniform(-2, W + 2, n_feat), rng.uniform(-2, H + 2, n_feat)], 1).astype(np.float32)   # sub-pixel positions
    obj_label = rng.integers(-1, 10, n_feat).astype(np.int32)
    perm = rng.permutation(n_feat)
    cuts = np.sort(rng.choice(np.arange(1, n_feat), n_obj, replace=False))
    inlier_sets = [perm[a:b] for a, b in zip(np.concatenate([[0], cuts[:-1]]), cuts)]
    a = rng.normal(0, 0.1, 3)
    th = np.linalg.norm(a); k = a / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    T = np.eye(4); T[:3, :3] = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx; T[:3, 3] = rng.normal(0, 1, 3)
    T = T.astype(np.float32)
    want = ol.renew_objects(cfg, ok, obj_label, inlier_sets, obj_stat, sem_pos, mod_label, tmp, depth, flow, mask, T)
    got = renew_objects_py(W, H, max_num_obj, (F(cam["fx"]), F(cam["fy"]), F(cam["cx"]), F(cam["cy"])), ok, obj_label, inlier_sets, obj_stat,
                           sem_pos, mod_label, tmp, depth, flow, mask, T)
    assert len(got[0]) == len(want[0]) and len(want[0]) > 50
    for g, w, name in zip(got[:7], want[:7], ("keys", "depth", "corres", "flow", "semantic label", "inlier id", "object label")):
        assert np.array_equal(g, w), name
    assert np.abs(got[7] - want[7]).max() <= 2e-5 * max(1.0, np.abs(want[7]).max())
    assert (want[6] == -2).any() and (want[5] >= 0).any() and ((want[5] == -1) & (want[6] > 0)).any()   # new / carried / topped-up
