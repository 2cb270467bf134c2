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
