"""CPU: the per-frame glue of the static path -- Tracking::GrabImageRGBD (src/Tracking.cc:283-421), Tracking::Initialization
(:1512-1580) and Tracking::Track (:1081-1160, 1310-1425) -- composed a second time, in Python and from the reference's text, out of
the oracle's separately pinned stages (depth pre-scale, ORB extraction, association, initial camera model, joint flow + pose
optimisation, feature renewal): the correspondences carried from frame to frame with their truncating depth look-ups, the 3-D
points handed to the PnP stage, the constant-velocity prior, the refined flow written back into the matches, the outliers removed
from the temporal match list, the motion model, and what is pushed into the Map.  The oracle tracker, run on the same frames,
returns the same poses and stores the same feature lists.  (The window optimisation only rewrites Map poses / points, never the
frame pose the next frame starts from -- src/Optimizer.cc:1056-1142 -- so the tracked poses do not depend on it.)"""
import numpy as np

import oracle_lib as ol
import synth

F = np.float32


def inv_pose(T):
    """Converter::toInvMatrix (src/Converter.cc:155-170): [R^T | -R^T t], float32 result of double sums"""
    T = np.asarray(T, np.float32)
    o = np.zeros((4, 4), np.float32)
    o[:3, :3] = T[:3, :3].T
    for r in range(3):
        o[r, 3] = F(sum(np.float64(-o[r, k]) * np.float64(T[k, 3]) for k in range(3)))
    o[3, 3] = 1
    return o


def mat_mul(A, B):
    """cv::Mat float product: double accumulation, one rounding"""
    o = np.zeros((4, 4), np.float32)
    for r in range(4):
        for c in range(4):
            s = np.float64(0)
            for k in range(4):
                s += np.float64(A[r, k]) * np.float64(B[k, c])
            o[r, c] = F(s)
    return o


def unproject_world(key, z, K, Twc):
    """Frame::UnprojectStereoStat (src/Frame.cc:706-735): camera-frame point in float32, then R*x + t"""
    fx, fy, cx, cy = K
    invfx, invfy = F(F(1) / fx), F(F(1) / fy)
    x = F(F(F(key[0] - cx) * z) * invfx)
    y = F(F(F(key[1] - cy) * z) * invfy)
    out = np.zeros(3, np.float32)
    for r in range(3):
        out[r] = F(F(np.float64(Twc[r, 0]) * np.float64(x) + np.float64(Twc[r, 1]) * np.float64(y) + np.float64(Twc[r, 2]) * np.float64(z)) + Twc[r, 3])
    return out


class PyTracker:
    def __init__(self, cam, cfg):
        self.cam, self.cfg = cam, cfg
        self.K = (F(cam["fx"]), F(cam["fy"]), F(cam["cx"]), F(cam["cy"]))
        self.p = ol.default_orb_params(cfg.orb.nfeatures)
        self.last = None
        self.velocity = None
        self.map = []        # per frame: (xy, depth, asso)

    def track(self, gray, depth_in, flow, mask):
        W, H = self.cam["width"], self.cam["height"]
        cfg = self.cfg
        depth = ol.depth_prep(depth_in, cfg.choose_data, cfg.depth_map_factor, cfg.bf)                      # :299-322
        cur = {}
        cur["kps"] = ol.orb_extract(gray, self.p)                                                           # Frame::Frame
        idx, cor, fl, dep = ol.frame_associate(cur["kps"], depth, flow, mask, cfg.th_depth_bg)              # Frame.cc:72-100
        cur["keys_tmp"] = np.stack([cur["kps"]["x"][idx], cur["kps"]["y"][idx]], 1).astype(np.float32)
        cur["corres"], cur["flow_next"], cur["depth_tmp"] = cor, fl, dep
        if self.last is None:                                                                               # Initialization
            cur["Tcw"] = np.eye(4, dtype=np.float32)
            self.map.append((cur["keys_tmp"], cur["depth_tmp"], None))
            cur["stat_keys"], cur["stat_depth"] = cur["keys_tmp"], cur["depth_tmp"]
            self.last = cur
            return cur["Tcw"]
        last = self.last
        # :369-389 the matches of this frame are the last frame's correspondences; depth at the truncated position, -1 if not > 0
        sk = last["corres"].copy()
        n = len(sk)
        sd = np.full(n, -1, np.float32)
        for i in range(n):
            v, u = int(sk[i, 1]), int(sk[i, 0])
            if 0 < u < W - 1 and 0 < v < H - 1:
                d = depth[v, u]
                if d > 0:
                    sd[i] = d
        # GetInitModelCam (:1914-1960): this frame's 2-D matches against the LAST frame's features lifted to the world with ITS depth
        Twl = inv_pose(last["Tcw"])
        valid = np.ones(n, np.int32)
        p3d = np.zeros((n, 3), np.float32)
        for i in range(n):
            z = last["stat_depth"][i]
            if z < 0:
                valid[i] = 0
                continue
            p3d[i] = unproject_world(last["stat_keys"][i], z, self.K, Twl)
        prior = mat_mul(self.velocity, last["Tcw"]) if self.velocity is not None else last["Tcw"]            # :1976
        T0, ids, winner, nr, nm = ol.init_model_cam(sk, p3d, valid, prior, self.K)
        TM = ids.copy()
        if cfg.b_joint:
            # PoseOptimizationFlow2Cam (:1133): the last frame's features, their flow and depth; refined flow replaces the match
            T1, fo, inl, ninl, _ = ol.poseopt_flow2cam(last["stat_keys"][TM], last["flow_next"][TM], last["stat_depth"][TM], T0, last["Tcw"], self.K)
            if len(TM) >= 3:
                for i, k in enumerate(TM):
                    if inl[i]:
                        sk[k, 0] = F(np.float64(last["stat_keys"][k, 0]) + np.float64(fo[i, 0]))
                        sk[k, 1] = F(np.float64(last["stat_keys"][k, 1]) + np.float64(fo[i, 1]))
                    else:
                        TM[i] = -1
        else:
            # PoseOptimizationNew (:1135, src/Optimizer.cc:2180-2334): reprojection of the last frame's world points, matches untouched
            T1, inl, _ = ol.pose_opt_proj(0, sk[TM], p3d[TM], T0, K=self.K)
            if len(TM) >= 3:
                TM[inl == 0] = -1
        cur["Tcw"] = T1
        self.velocity = mat_mul(T1, Twl)                                                                    # :1142-1148
        # RenewFrameInfo (:1320) and the hand-over (:1336-1340)
        keys, corres, fnext, inlier, dtmp, p3 = ol.renew_static(cfg, TM, sk, cur["kps"], depth, flow, mask, T1)
        cur["stat_keys"], cur["stat_depth"], cur["corres"], cur["flow_next"] = keys, dtmp, corres, fnext
        self.map.append((keys, dtmp, inlier))
        self.last = cur
        return T1


def test_static_sequence_composed_in_python_equals_the_oracle_tracker():
    cam = synth.KITTI
    sc = synth.Scene(cam=cam, seed=4242, flow_noise=0.1, depth_noise=0.01)
    cfg = ol.track_config(cam)
    tr = ol.OracleTracker(cfg)
    py = PyTracker(cam, cfg)
    n = 8
    for k in range(n):
        f = sc.frame(k)
        g, d, fl, m = f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy()
        T_or, st, rc = tr.track(g, d, fl, m)
        assert rc == 0
        T_py = py.track(g, d, fl, m)
        assert np.array_equal(T_py, T_or), k                      # the pose returned to the caller, bit for bit
        xy, dep, p3, asso = tr.static_features(k)
        kx, kd, ka = py.map[k]
        assert np.array_equal(xy, kx) and np.array_equal(dep, kd), k
        if k > 0:
            assert np.array_equal(asso, ka), k
            assert st["n_static"] == len(kx) and len(kx) >= 900
    tr.close()


# ---------------------------------------------------------------------------------------------------------------------------
# the object path on top: UpdateMask, object samples and their carry-over (src/Tracking.cc:343-364, 391-421), GetSceneFlowObj
# (:1582-1668), DynObjTracking (:1670-1912, the Python restatement of test_dynobj_independent), the per-object loop of Track
# (:1179-1308) with GetInitModelObj (:2030-2162) and PoseOptimizationFlow2 (src/Optimizer.cc:3037), the object half of
# RenewFrameInfo and the Map pushes (:1345-1422)
# ---------------------------------------------------------------------------------------------------------------------------
from test_dynobj_independent import dyn_obj_tracking_py  # noqa: E402


class PyDynTracker(PyTracker):
    def __init__(self, cam, cfg):
        super().__init__(cam, cfg)
        self.seg_last = self.flow_last = None
        self.max_id = 1
        self.f_id = 0
        self.dyn_map = []     # per frame: (xy, depth, p3, asso, label)
        self.obj_map = []     # per frame >= 1: (label, semantic label, motion, centre)

    def world(self, key, z, Tcw):
        return unproject_world(key, z, self.K, inv_pose(Tcw))

    def track(self, gray, depth_in, flow, mask):
        W, H = self.cam["width"], self.cam["height"]
        cfg = self.cfg
        last = self.last
        depth = ol.depth_prep(depth_in, cfg.choose_data, cfg.depth_map_factor, cfg.bf)
        if last is not None and len(last["obj_sem"]) and self.seg_last is not None:            # UpdateMask, :343-364
            mask, _, _ = ol.update_mask(last["obj_sem"], last["obj_corres"], self.seg_last, self.flow_last, mask)
        T = super().track(gray, depth_in, flow, mask)
        cur = self.last
        tk, tc, tf, td, ts = ol.frame_sample_objects(depth, flow, mask, cfg.th_depth_obj)       # Frame.cc:184-211
        tmp = dict(keys=tk, corres=tc, flow=tf, depth=td, sem=ts)
        if last is None:
            cur.update(obj_keys=tk, obj_corres=tc, obj_flow_next=tf, obj_depth=td, obj_sem=ts, obj_label=np.full(len(tk), -2, np.int32),
                       mod_label=np.zeros(0, np.int32), sem_pos=np.zeros(0, np.int32), obj_stat=np.zeros(0, np.int32), obj_mod=[])
            fx, fy, cx, cy = self.K
            self.dyn_map.append((tk, td, None, None, None))
        else:
            # :403-421 object features of this frame = the last frame's correspondences, fresh depth / label at the truncated position
            ok = last["obj_corres"].copy()
            n = len(ok)
            od = np.zeros(n, np.float32); osem = np.zeros(n, np.int32)
            for i in range(n):
                u, v = int(ok[i, 0]), int(ok[i, 1])
                if 0 < u < W - 1 and 0 < v < H - 1 and 0 < depth[v, u] < cfg.th_depth_obj:
                    od[i] = depth[v, u]; osem[i] = mask[v, u]
                else:
                    od[i] = F(0.1); osem[i] = 0
            lab = np.full(n, -2, np.int32)
            inlier_sets, stat, mods, cents, mod_label, sem_pos = [], [], [], [], np.zeros(0, np.int32), np.zeros(0, np.int32)
            if n:
                # GetSceneFlowObj: world displacement of every feature with a semantic label in both frames
                flow3 = np.zeros((n, 3), np.float32)
                for i in range(n):
                    if osem[i] <= 0 or last["obj_sem"][i] <= 0:
                        lab[i] = -1
                        continue
                    flow3[i] = self.world(ok[i], od[i], T) - self.world(last["obj_keys"][i], last["obj_depth"][i], last["Tcw"])
                lab, self.max_id, mod_label, sem_pos, ids = dyn_obj_tracking_py(
                    W, H, cfg.sf_mg_thres, cfg.sf_ds_thres, cfg.th_depth_obj, osem, lab, ok, od, flow3, last["obj_sem"], last["sem_pos"],
                    last["obj_stat"], last["mod_label"], self.f_id, self.max_id)
                Twc = inv_pose(T)
                for o, oid in enumerate(ids):                                                    # :1179-1308
                    p3 = np.stack([self.world(last["obj_keys"][i], last["obj_depth"][i], last["Tcw"]) for i in oid])
                    c = np.zeros(3, np.float32)
                    for p in p3:
                        c = (c + p).astype(np.float32)
                    centre = (c * F(1.0 / np.float64(len(oid)))).astype(np.float32)
                    pre = [k for k in range(len(last["mod_label"])) if last["mod_label"][k] == mod_label[o]]     # GetInitModelObj
                    if pre:
                        T0, sub, _, _, _ = ol.init_model_cam(ok[oid], p3, None, mat_mul(T, last["obj_mod"][pre[0]]), self.K)
                    else:
                        T0, sub, _, _, _ = ol.init_model_cam(ok[oid], p3, None, T, self.K, no_motion_model=1)
                    keep = np.zeros(len(oid), bool); keep[sub] = True
                    lab[oid[~keep]] = -1
                    in_ids = oid[sub]
                    if len(in_ids) < 50:
                        stat.append(0); mods.append(np.eye(4, dtype=np.float32)); cents.append(np.zeros(3, np.float32)); inlier_sets.append(in_ids)
                        continue
                    Tx, fo, inl, _, _ = ol.poseopt_flow2cam(last["obj_keys"][in_ids], last["obj_flow_next"][in_ids], last["obj_depth"][in_ids], T0,
                                                            last["Tcw"], self.K, info_prior=0.5, rounds=1, its=200)
                    mods.append(mat_mul(Twc, Tx))                                                # vObjMod = inv(Tcw) * Obj_X
                    good = []
                    for k, i in enumerate(in_ids):
                        if inl[k]:
                            ok[i, 0] = F(np.float64(last["obj_keys"][i, 0]) + np.float64(fo[k, 0]))
                            ok[i, 1] = F(np.float64(last["obj_keys"][i, 1]) + np.float64(fo[k, 1]))
                            good.append(i)
                        else:
                            lab[i] = -1
                    stat.append(1); cents.append(centre); inlier_sets.append(np.array(good, np.int32))
            nk, nd, ncor, nfl, nsem, ninl, nlab, np3 = ol.renew_objects(cfg, ok, lab, inlier_sets, stat, sem_pos, mod_label, tmp, depth, flow, mask, T)
            cur.update(obj_keys=nk, obj_depth=nd, obj_corres=ncor, obj_flow_next=nfl, obj_sem=nsem, obj_label=nlab, mod_label=mod_label,
                       sem_pos=sem_pos, obj_stat=np.array(stat, np.int32), obj_mod=mods)
            self.dyn_map.append((nk, nd, np3, ninl, nlab))
            okk = [o for o in range(len(stat)) if stat[o]]
            self.obj_map.append((mod_label[okk], sem_pos[okk], [mods[o] for o in okk], [cents[o] for o in okk]))
        if len(cur["obj_sem"]):
            self.seg_last, self.flow_last = mask.copy(), flow.copy()
        else:
            self.seg_last = self.flow_last = None
        self.f_id += 1
        return T


def test_dynamic_sequence_composed_in_python_equals_the_oracle_tracker():
    cam = synth.KITTI
    sc = synth.Scene(cam=cam, seed=1234, flow_noise=0.05, depth_noise=0.005, n_objects=5, drop_mask=[(3, 2)])
    cfg = ol.track_config(cam)
    tr = ol.OracleTracker(cfg)
    py = PyDynTracker(cam, cfg)
    n = 6
    for k in range(n):
        f = sc.frame(k)
        g, d, fl, m = f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy()
        T_or, st, rc = tr.track(g, d, fl, m)
        assert rc == 0
        T_py = py.track(g, d, fl, m)
        assert np.array_equal(T_py, T_or), k
        xy, dep, p3, asso, lab = tr.dynamic_features(k)
        kx, kd, kp3, ka, kl = py.dyn_map[k]
        assert len(xy) > 2000 and np.array_equal(xy, kx) and np.array_equal(dep, kd), k
        if k > 0:
            assert np.array_equal(asso, ka) and np.array_equal(lab, kl) and np.array_equal(p3, kp3), k
            olab, osem, omot, ocen = tr.objects(k)
            plab, psem, pmot, pcen = py.obj_map[k - 1]
            assert len(olab) == 5 and np.array_equal(olab, plab) and np.array_equal(osem, psem), k
            assert np.array_equal(omot, np.stack(pmot)) and np.array_equal(ocen, np.stack(pcen)), k
            assert st["n_masks_recovered"] == (1 if k == 3 else 0)
    tr.close()


# ---------------------------------------------------------------------------------------------------------------------------
# the window optimisation chained over the sequence: what Tracking::Track pushes into the Map (:1345-1422), GetStaticTrack, the
# graph Optimizer::PartialBatchOptimization builds from it (src/Optimizer.cc:43-100, 218-362) and what it writes back
# (:1056-1142) -- every window starts from the poses, relative motions and points the windows before it left in the Map
# ---------------------------------------------------------------------------------------------------------------------------
from test_tracklets_independent import get_static_track  # noqa: E402


def camera_point(key, z, K):
    """Optimizer::Get3DinCamera (src/Optimizer.cc:3277-3294)"""
    fx, fy, cx, cy = K
    invfx, invfy = F(F(1) / fx), F(F(1) / fy)
    return np.array([F(F(F(key[0] - cx) * z) * invfx), F(F(F(key[1] - cy) * z) * invfy), z], np.float32)


class PyMapTracker(PyTracker):
    def __init__(self, cam, cfg):
        super().__init__(cam, cfg)
        self.poses, self.rel, self.p3 = [], [], []
        self.ba = []

    def track(self, gray, depth_in, flow, mask):
        first = self.last is None
        T = super().track(gray, depth_in, flow, mask)
        xy, dep, asso = self.map[-1]
        if first:                                                          # Initialization: camera-frame points, identity pose
            self.p3.append(np.stack([camera_point(xy[i], dep[i], self.K) for i in range(len(xy))]))
            self.poses.append(np.eye(4, dtype=np.float32))
            return T
        Twc = inv_pose(T)
        self.p3.append(np.stack([unproject_world(xy[i], dep[i], self.K, Twc) for i in range(len(xy))]))   # RenewFrameInfo (4)
        self.poses.append(Twc)                                             # vmCameraPose: Twc
        self.rel.append(inv_pose(self.velocity))                           # vmRigidMotion[.][0]: inverse of the motion model
        self.partial_batch(min(len(self.poses) - 1, self.cfg.window_size))
        return T

    def partial_batch(self, window):
        N = len(self.poses)
        tracks = get_static_track([m[2].tolist() for m in self.map[1:]])
        label = [[-1] * len(m[0]) for m in self.map]
        for t, tr in enumerate(tracks):
            if len(tr) >= 3:
                for f, j in tr:
                    label[f][j] = t
        mark = [[-1] * len(m[0]) for m in self.map]
        start = N - window
        pts, owner, op, opt, oxyz = [], [], [], [], []
        for i in range(start, N):
            xy, dep, _ = self.map[i]
            for j in range(len(xy)):
                t = label[i][j]
                if t == -1:
                    continue
                pos = tracks[t].index((i, j))
                if pos == 0:
                    pid = len(pts)
                    pts.append(self.p3[i][j]); owner.append((i, j))
                else:
                    pf, pj = tracks[t][pos - 1]
                    pid = mark[pf][pj]
                    if pid == -1:
                        continue
                mark[i][j] = pid
                op.append(i - start); opt.append(pid); oxyz.append(camera_point(xy[j], dep[j], self.K))
        poses, rel, points, its, st = ol.ba_partial(np.stack(self.poses[start:]).reshape(window, 16),
                                                    np.stack(self.rel[start:]).reshape(-1, 16) if window > 1 else np.zeros((0, 16), np.float32),
                                                    np.array(pts, np.float32).reshape(-1, 3), op, opt, np.array(oxyz, np.float32).reshape(-1, 3))
        self.ba.append((len(pts), len(op), its, st.total_trials))
        for i in range(start, N):                                          # write-back, :1056-1142
            self.poses[i] = poses[i - start].reshape(4, 4)
            if i > start:
                self.rel[i - 1] = rel[i - start - 1].reshape(4, 4)
            for j in range(len(mark[i])):
                if mark[i][j] != -1:
                    self.p3[i][j] = points[mark[i][j]]


def test_window_optimisation_chain_composed_in_python_equals_the_oracle_tracker():
    cam = synth.SMALL
    W = 6                                                                  # the window slides from frame 7 on
    sc = synth.Scene(cam=cam, seed=77, flow_noise=0.1, depth_noise=0.01)
    cfg = ol.track_config(cam, nfeatures=1200, max_track_bg=400, window=W)
    tr = ol.OracleTracker(cfg)
    py = PyMapTracker(cam, cfg)
    n = 12
    for k in range(n):
        f = sc.frame(k)
        g, d, fl, m = f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy()
        T_or, st, rc = tr.track(g, d, fl, m)
        assert rc == 0
        assert np.array_equal(py.track(g, d, fl, m), T_or), k
        if k > 0:
            assert (st["ba_points"], st["ba_obs"], st["ba_iterations"], st["ba_trials"]) == py.ba[-1], k
        assert np.array_equal(tr.map_poses(), np.stack(py.poses)), k                      # every Map pose after this frame's window
        for j in range(k + 1):
            assert np.array_equal(tr.static_features(j)[2], py.p3[j]), (k, j)             # and every Map point
    assert py.ba[-1][0] > 100 and py.ba[-1][2] >= 2
    tr.close()


def test_reprojection_only_branch_composed_in_python_equals_the_oracle_tracker():
    """Tracking::bJoint = false: PoseOptimizationNew instead of the joint flow + pose optimisation (src/Tracking.cc:1133-1136)"""
    cam = synth.SMALL
    sc = synth.Scene(cam=cam, seed=31, flow_noise=0.1, depth_noise=0.01)
    cfg = ol.track_config(cam, nfeatures=1200, max_track_bg=400, b_joint=0)
    tr = ol.OracleTracker(cfg)
    py = PyTracker(cam, cfg)
    for k in range(6):
        f = sc.frame(k)
        g, d, fl, m = f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy()
        T_or, st, rc = tr.track(g, d, fl, m)
        assert rc == 0
        assert np.array_equal(py.track(g, d, fl, m), T_or), k
        xy, dep, p3, asso = tr.static_features(k)
        assert np.array_equal(xy, py.map[k][0]) and np.array_equal(dep, py.map[k][1]), k
        if k > 0:
            assert np.array_equal(asso, py.map[k][2]), k
    tr.close()
