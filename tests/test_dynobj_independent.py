"""CPU: Tracking::DynObjTracking (src/Tracking.cc:1670-1912) restated a second time, in Python and from the reference's text,
and run (a) on the inputs of every call the oracle tracker made on a five-object sequence and (b) on randomised frame states that
reach the branches that sequence never takes (objects on the image boundary, static / far / small objects, ids lost and re-used,
invalid previous motions, ties of the majority vote).  The product's object lists, labels and tracking ids are compared with the
oracle's on the GPU (tests/test_dyn_gpu.py, tests/test_long_sequence_gpu.py); this file ties the oracle to an independent reading."""
import numpy as np
import pytest

import oracle_lib as ol
import synth

F = np.float32


def dyn_obj_tracking_py(W, H, sf_mg, sf_ds, th_depth_obj, sem, lab, key_xy, depth, flow3, last_sem, last_sem_pos, last_stat, last_mod,
                        f_id, max_id):
    lab = lab.copy()
    # :1677-1703 unique semantic labels (ascending) and the features of each that are not outliers, in feature order
    uni = sorted(set(int(s) for s in sem))
    posi = {u: [i for i in range(len(sem)) if sem[i] == u and lab[i] != -1] for u in uni}
    # :1705-1733 objects whose features lie mostly (> 50 %) in the 10-row / 20-column boundary band are discarded
    kept = []
    for u in uni:
        idx = posi[u]
        cnt = F(0)
        for i in idx:
            x, y = key_xy[i]
            if y < 10 or y > H - 10 or x < 20 or x > W - 20:
                cnt = F(cnt + F(1))
        # an empty list divides 0 by 0 in the reference: NaN > 0.5 is false, the (empty) object goes on
        with np.errstate(invalid="ignore", divide="ignore"):
            frac = F(cnt) / F(len(idx))
        if frac > F(0.5):
            for i in idx:
                lab[i] = -1
        else:
            kept.append((u, idx))
    # :1735-1805 scene-flow statistics per object: static -> label 0, far away or fewer than 150 features -> label -1
    objs = []
    for u, idx in kept:
        dsum, slow = F(0), F(0)
        for i in idx:
            dsum = F(dsum + depth[i])
            nrm = F(np.sqrt(F(F(flow3[i][0] * flow3[i][0]) + F(flow3[i][2] * flow3[i][2]))))
            if nrm < F(sf_mg):
                slow = F(slow + F(1))
        with np.errstate(invalid="ignore", divide="ignore"):
            static = F(slow) / F(len(idx)) > F(sf_ds)
            far = F(dsum) / F(len(idx)) > F(th_depth_obj)
        if static:
            for i in idx:
                lab[i] = 0
        elif far or len(idx) < 150:
            for i in idx:
                lab[i] = -1
        else:
            objs.append((u, idx))
    # :1841-1895 tracking ids: majority semantic label of the object's features in the LAST frame; an object of the last frame
    # with that semantic label and a valid motion hands its id on, otherwise a fresh id
    if f_id == 1:
        max_id = 1
    mod = []
    for u, idx in objs:
        votes = {}
        for i in idx:
            votes[int(last_sem[i])] = votes.get(int(last_sem[i]), 0) + 1
        best = max(sorted(votes), key=lambda k: votes[k])      # most votes; among equals the smallest label (std::map order)
        new_id = None
        if max_id != 1:
            for k in range(len(last_sem_pos)):
                if last_sem_pos[k] == best and last_stat[k]:
                    new_id = int(last_mod[k])
                    break
        if new_id is None:
            new_id = max_id
            max_id += 1
        for i in idx:
            lab[i] = new_id
        mod.append(new_id)
    return lab, max_id, np.array(mod, np.int32), np.array([u for u, _ in objs], np.int32), [np.array(i, np.int32) for _, i in objs]


def same(a, b):
    la, ma, moda, sema, idsa = a
    lb, mb, modb, semb, idsb = b
    assert np.array_equal(la, lb)
    assert ma == mb
    assert np.array_equal(moda, modb) and np.array_equal(sema, semb)
    assert len(idsa) == len(idsb)
    for x, y in zip(idsa, idsb):
        assert np.array_equal(x, y)


def test_recorded_calls_of_a_five_object_sequence():
    cam = synth.KITTI
    sc = synth.Scene(cam=cam, seed=1234, flow_noise=0.05, depth_noise=0.005, n_objects=5, drop_mask=[(3, 2)])
    cfg = ol.track_config(cam)
    tr = ol.OracleTracker(cfg)
    tr.dyn_log_enable()
    for k in range(7):
        f = sc.frame(k)
        tr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
    log = tr.dyn_log()
    tr.close()
    assert len(log) == 6
    for r in log:
        at, ids = 0, []
        for n in r["out_len"]:
            ids.append(r["out_ids"][at:at + n]); at += n
        got = dyn_obj_tracking_py(cam["width"], cam["height"], cfg.sf_mg_thres, cfg.sf_ds_thres, cfg.th_depth_obj, r["sem"], r["lab_before"],
                                  r["key_xy"], r["depth"], r["flow3"], r["last_sem"], r["last_sem_pos"], r["last_stat"], r["last_mod"],
                                  r["f_id"], r["max_id_before"])
        same(got, (r["lab_after"], r["max_id_after"], r["out_mod"], r["out_sem_pos"], ids))
        assert len(ids) == 5 and len(r["sem"]) > 2000


@pytest.mark.parametrize("seed", range(12))
def test_randomised_frame_states(seed):
    rng = np.random.default_rng(seed)
    cam = synth.KITTI
    W, H = cam["width"], cam["height"]
    cfg = ol.track_config(cam)
    nobj = int(rng.integers(1, 9))
    sem_ids = rng.choice(np.arange(1, 30), nobj, replace=False)
    sem, key, dep, fl, lsem = [], [], [], [], []
    for s in sem_ids:
        n = int(rng.choice([5, 120, 149, 150, 151, 400]))
        kind = rng.integers(0, 5)          # 0 moving, 1 static, 2 far, 3 on the boundary, 4 half slow (near the 30 % threshold)
        cx, cy = rng.uniform(100, W - 100), rng.uniform(60, H - 60)
        xy = np.stack([rng.uniform(cx - 60, cx + 60, n), rng.uniform(cy - 40, cy + 40, n)], 1)
        if kind == 3:
            xy[:, 0] = rng.uniform(0, 40, n)           # about half inside the 20-column band
        d = rng.uniform(5, 20, n) if kind != 2 else rng.uniform(20, 31, n)   # mean around the 25 m gate for "far"
        speed = {0: 0.8, 1: 0.02, 2: 0.8, 3: 0.8, 4: 0.8}[int(kind)]
        f3 = rng.normal(0, 0.01, (n, 3)) + [0.0, 0.0, speed]
        if kind == 4:
            slow = rng.random(n) < 0.3
            f3[slow] = rng.normal(0, 0.01, (int(slow.sum()), 3))
        f3[:, 1] = rng.normal(0, 5.0, n)              # the y component must not matter
        ls = np.full(n, s)
        flip = rng.random(n) < 0.2
        ls[flip] = rng.integers(0, 30, int(flip.sum()))     # some features sat on another label in the last frame
        if rng.random() < 0.25:                              # a tie of the majority vote
            ls[: n // 2] = s; ls[n // 2: 2 * (n // 2)] = int(s) + 1 if int(s) > 1 else int(s) - 1 + 2
        sem += [s] * n; key.append(xy); dep.append(d); fl.append(f3); lsem.append(ls)
    sem = np.array(sem, np.int32)
    perm = rng.permutation(len(sem))                         # features of different objects interleaved
    sem = sem[perm]
    key = np.concatenate(key)[perm].astype(np.float32); dep = np.concatenate(dep)[perm].astype(np.float32)
    fl = np.concatenate(fl)[perm].astype(np.float32); lsem = np.concatenate(lsem)[perm].astype(np.int32)
    lab = np.where(rng.random(len(sem)) < 0.1, -1, 0).astype(np.int32)   # outliers of the previous frame
    nlast = int(rng.integers(0, 8))
    last_sem_pos = rng.choice(np.arange(1, 30), nlast, replace=False).astype(np.int32) if nlast else np.zeros(0, np.int32)
    if nlast and rng.random() < 0.8:
        last_sem_pos[: min(nlast, nobj)] = sem_ids[: min(nlast, nobj)]
    last_stat = (rng.random(nlast) < 0.7).astype(np.int32)
    last_mod = (rng.permutation(20)[:nlast] + 1).astype(np.int32)
    f_id = int(rng.choice([1, 2, 9]))
    max_id = int(rng.choice([1, 2, 7, 25]))
    want = ol.dyn_obj_tracking(cfg, sem, lab, key, dep, fl, lsem, last_sem_pos, last_stat, last_mod, f_id, max_id)
    got = dyn_obj_tracking_py(W, H, cfg.sf_mg_thres, cfg.sf_ds_thres, cfg.th_depth_obj, sem, lab, key, dep, fl, lsem, last_sem_pos, last_stat,
                              last_mod, f_id, max_id)
    same(got, want)


def test_randomised_states_reach_every_branch():
    """the generator above produces, over its seeds, every outcome: kept objects, static (0), discarded (-1), re-used and fresh ids"""
    cam = synth.KITTI
    cfg = ol.track_config(cam)
    seen = set()
    for seed in range(12):
        rng = np.random.default_rng(1000 + seed)
        n = 300
        sem = np.repeat([3, 4], n).astype(np.int32)
        key = np.stack([rng.uniform(200, 400, 2 * n), rng.uniform(100, 200, 2 * n)], 1).astype(np.float32)
        dep = np.full(2 * n, 10, np.float32)
        fl = np.zeros((2 * n, 3), np.float32); fl[:n, 2] = 1.0          # object 3 moves, object 4 is static
        lab = np.zeros(2 * n, np.int32)
        out = ol.dyn_obj_tracking(cfg, sem, lab, key, dep, fl, sem, np.array([3], np.int32), np.array([seed % 2], np.int32),
                                  np.array([5], np.int32), 4, 9)
        assert (out[0][n:] == 0).all()                                   # static object -> background label
        assert out[2].tolist() == ([5] if seed % 2 else [9])             # id handed on only from a valid previous motion
        seen.add(int(out[2][0]))
    assert seen == {5, 9}
