"""CPU: hand-checkable cases for the oracle's restatement of Tracking::UpdateMask (src/Tracking.cc:3291-3357) and the
stride-4 object sampling of Frame::Frame (src/Frame.cc:184-211)."""
import numpy as np

import oracle_lib as ol


def test_update_mask_hand_case():
    H, W = 40, 60
    mask_last = np.zeros((H, W), np.int32)
    mask_last[8:28, 10:30] = 5                      # 20 x 20 block of label 5
    flow_last = np.zeros((H, W, 2), np.float32)
    flow_last[..., 0] = 3.9                         # truncates to +3
    flow_last[..., 1] = -1.7                        # truncates to -1 (toward zero)
    ys, xs = np.nonzero(mask_last == 5)
    sem = np.full(len(ys), 5, np.int32)
    corres = np.stack([xs + 3.9, ys - 1.7], 1).astype(np.float32)
    # (1) mask lost in the new frame -> recovered at the truncated-flow positions
    out, uniq, rec = ol.update_mask(sem, corres, mask_last, flow_last, np.zeros((H, W), np.int32))
    want = np.zeros((H, W), np.int32)
    want[7:27, 13:33] = 5
    assert list(uniq) == [5] and list(rec) == [1] and np.array_equal(out, want)
    # (2) mask still there (any non-zero majority) -> untouched
    cur = np.zeros((H, W), np.int32); cur[5:30, 10:36] = 9
    out, _, rec = ol.update_mask(sem, corres, mask_last, flow_last, cur)
    assert list(rec) == [0] and np.array_equal(out, cur)
    # (3) fewer than 100 votes -> untouched even though the mask is lost
    out, _, rec = ol.update_mask(sem[:99], corres[:99], mask_last, flow_last, np.zeros((H, W), np.int32))
    assert list(rec) == [0] and not out.any()
    # (4) a tie between label 0 and another label goes to the smaller one (0): recovered
    cur = np.zeros((H, W), np.int32)
    votes = np.stack([corres[:, 0].astype(np.int32), corres[:, 1].astype(np.int32)], 1)
    half = votes[: len(votes) // 2]
    cur[half[:, 1], half[:, 0]] = 4
    out, _, rec = ol.update_mask(sem, corres, mask_last, flow_last, cur)
    assert list(rec) == [1]


def test_object_sampling_order_and_gates():
    H, W = 16, 24
    depth = np.full((H, W), 5.0, np.float32)
    mask = np.zeros((H, W), np.int32); mask[4:12, 8:20] = 2
    flow = np.zeros((H, W, 2), np.float32); flow[..., 0] = 1.5; flow[..., 1] = 0.5
    depth[8, 12] = 30.0      # beyond ThDepthOBJ
    flow[4, 16] = (100.0, 0.0)  # leaves the image
    keys, cor, fl, dep, lab = ol.frame_sample_objects(depth, flow, mask, 25.0)
    want = [(x, y) for y in range(0, H, 4) for x in range(0, W, 4) if mask[y, x] and (x, y) not in ((12, 8), (16, 4))]
    assert [tuple(k) for k in keys.astype(int)] == want and (lab == 2).all()
    assert np.array_equal(cor, keys + np.float32([1.5, 0.5]))


def test_association_and_sampling_against_pure_python_restatements():
    """Frame::Frame's static association (src/Frame.cc:72-100) and stride-4 object sampling (:184-211) restated line by line in
    pure Python (float32 arithmetic, the same comparison order) on random maps: identical lists in identical order."""
    rng = np.random.default_rng(5)
    H, W = 61, 83
    f32 = np.float32
    for trial in range(3):
        depth = rng.uniform(-2, 60, (H, W)).astype(f32)
        mask = (rng.integers(0, 4, (H, W)) * (rng.random((H, W)) < 0.4)).astype(np.int32)
        flow = rng.normal(0, 6, (H, W, 2)).astype(f32)
        flow[rng.random((H, W)) < 0.1] = 0
        n = 400
        xy = np.stack([rng.uniform(0, W - 1, n), rng.uniform(0, H - 1, n)], 1).astype(f32)
        kps = np.zeros(n, ol.KP_DTYPE)
        kps["x"], kps["y"] = xy[:, 0], xy[:, 1]
        th = 40.0
        idx, cor, fl, dep = ol.frame_associate(kps, depth, flow, mask, th)
        want = []
        for i in range(n):
            x, y = int(xy[i, 0]), int(xy[i, 1])
            if mask[y, x] != 0:
                continue
            if depth[y, x] > f32(th) or depth[y, x] <= 0:
                continue
            fx, fy = flow[y, x]
            if fx != 0 and fy != 0 and f32(xy[i, 0] + fx) < W and f32(xy[i, 1] + fy) < H and xy[i, 0] < W and xy[i, 1] < H:
                want.append((i, f32(xy[i, 0] + fx), f32(xy[i, 1] + fy), fx, fy))
        assert list(idx) == [w[0] for w in want]
        assert np.array_equal(cor, np.array([[w[1], w[2]] for w in want], f32)) and np.array_equal(fl, np.array([[w[3], w[4]] for w in want], f32))
        keys, ocor, ofl, odep, lab = ol.frame_sample_objects(depth, flow, mask, 25.0)
        wk = []
        for i in range(0, H, 4):
            for j in range(0, W, 4):
                if mask[i, j] != 0 and depth[i, j] < f32(25.0) and depth[i, j] > 0:
                    fx, fy = flow[i, j]
                    if f32(j + fx) < W and f32(j + fx) > 0 and f32(i + fy) < H and f32(i + fy) > 0:
                        wk.append((j, i, f32(j + fx), f32(i + fy), depth[i, j], mask[i, j]))
        assert [tuple(k) for k in keys.astype(int)] == [(w[0], w[1]) for w in wk]
        assert np.array_equal(ocor, np.array([[w[2], w[3]] for w in wk], f32)) and np.array_equal(odep, np.array([w[4] for w in wk], f32))
        assert np.array_equal(lab, np.array([w[5] for w in wk], np.int32))


def test_update_mask_against_a_pure_python_restatement():
    """Tracking::UpdateMask (src/Tracking.cc:3291-3357) line by line in pure Python on random scenes: labels processed in sorted
    order, votes read from the mask AS MODIFIED by the labels before, int truncation of correspondences and flows, >= 100 votes,
    majority with ties to the smaller label, in-place forward warp in raster order"""
    rng = np.random.default_rng(11)
    H, W = 48, 72
    for trial in range(6):
        mask_last = np.zeros((H, W), np.int32)
        labels = [2, 5, 7]
        boxes = [(4, 4, 18, 20), (20, 30, 40, 52), (8, 50, 24, 68)]
        for lab, (y0, x0, y1, x1) in zip(labels, boxes):
            mask_last[y0:y1, x0:x1] = lab
        flow_last = rng.normal(0, 3, (H, W, 2)).astype(np.float32)
        flow_last[..., 0] += rng.uniform(-4, 4); flow_last[..., 1] += rng.uniform(-3, 3)
        ys, xs = np.nonzero(mask_last)
        pick = rng.random(len(ys)) < 0.9
        ys, xs = ys[pick], xs[pick]
        order = rng.permutation(len(ys))
        ys, xs = ys[order], xs[order]
        sem = mask_last[ys, xs].astype(np.int32)
        corres = np.stack([xs + flow_last[ys, xs, 0], ys + flow_last[ys, xs, 1]], 1).astype(np.float32)
        mask_cur = np.zeros((H, W), np.int32)
        for lab, (y0, x0, y1, x1) in zip(labels, boxes):
            if rng.random() < 0.5:     # this label survives in the new frame (shifted), otherwise it is lost
                dy, dx = int(rng.integers(-3, 4)), int(rng.integers(-3, 4))
                mask_cur[max(y0 + dy, 0):y1 + dy, max(x0 + dx, 0):x1 + dx] = lab
        got, uniq, rec = ol.update_mask(sem, corres, mask_last, flow_last, mask_cur.copy())
        # ---- restatement
        seg = mask_cur.copy()
        uni = sorted(set(sem.tolist()))
        want_rec = []
        for lab in uni:
            votes = []
            for i in np.nonzero(sem == lab)[0]:
                u, v = int(corres[i, 0]), int(corres[i, 1])
                if 0 < u < W and 0 < v < H:
                    votes.append(int(seg[v, u]))
            r = 0
            if len(votes) >= 100:
                cnt = {}
                for k in votes:
                    cnt[k] = cnt.get(k, 0) + 1
                best = sorted(cnt.items(), key=lambda kv: (-kv[1], kv[0]))[0][0]
                if best == 0:
                    r = 1
                    for j in range(H):
                        for k in range(W):
                            if mask_last[j, k] == lab:
                                fx, fy = int(flow_last[j, k, 0]), int(flow_last[j, k, 1])
                                if 0 < k + fx < W and 0 < j + fy < H:
                                    seg[j + fy, k + fx] = lab
            want_rec.append(r)
        assert list(uniq) == uni and list(rec) == want_rec, (trial, list(rec), want_rec)
        assert np.array_equal(got, seg), trial


def test_depth_pre_scale_against_numpy():
    """Tracking::GrabImageRGBD's in-place depth pre-scale (src/Tracking.cc:299-322): float32 operations in the reference's order"""
    rng = np.random.default_rng(3)
    d = rng.uniform(-50, 9000, (37, 53)).astype(np.float32)
    d[rng.random(d.shape) < 0.05] = 0
    f32 = np.float32
    factor, bf, ms = f32(256.0), f32(386.1448), f32(0.93)
    with np.errstate(divide="ignore"):
        want = {1: np.where(d < 0, f32(0), d / factor), 2: np.where(d < 0, f32(0), bf / (d / factor)),
                3: np.where(d < 0, f32(0), (ms * bf) / (d / factor))}
    for mode in (1, 2, 3):
        got = ol.depth_prep(d, mode, float(factor), float(bf), float(ms))
        assert got.dtype == np.float32 and np.array_equal(got, want[mode].astype(np.float32)), mode
