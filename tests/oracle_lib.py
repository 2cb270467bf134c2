"""ctypes binding of oracle/liboracle.so (the CPU checker).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])


def default_orb_params(nfeatures=2500):
    return OrbParams(nfeatures, 1.2, 8, 20, 7)


def lib():
    global _LIB
    if _LIB is None:
        path = os.environ.get("VIDO_ORACLE_LIB") or os.path.join(ROOT, "oracle", "liboracle.so")   # bench.py: the -march=native build
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _LIB = C.CDLL(path)
        _LIB.vo_fast_atan2.restype = C.c_float
        _LIB.vo_fast_atan2.argtypes = [C.c_float, C.c_float]
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def level_sizes(W, H, p):
    w = np.zeros(p.nlevels, np.int32); h = np.zeros(p.nlevels, np.int32); s = np.zeros(p.nlevels, np.float32)
    lib().vo_orb_level_sizes(W, H, C.byref(p), _p(w), _p(h), _p(s))
    return w, h, s


def level_quotas(p):
    q = np.zeros(p.nlevels, np.int32)
    lib().vo_orb_level_quotas(C.byref(p), _p(q))
    return q


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().vo_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def fast_roi(img, thr, cap=1 << 20):
    img = np.ascontiguousarray(img, np.uint8)
    xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
    n = lib().vo_fast_roi(_p(img), img.shape[1], img.shape[0], img.shape[1], thr, _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def fast_atan2(y, x):
    return lib().vo_fast_atan2(float(y), float(x))


def level_candidates(img, p, cap=1 << 20):
    img = np.ascontiguousarray(img, np.uint8)
    xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
    n = lib().vo_orb_level_candidates(_p(img), img.shape[1], img.shape[0], img.shape[1], C.byref(p),
                                      _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def orb_pyramid(gray, p):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    w, h, _ = level_sizes(W, H, p)
    total = int((w.astype(np.int64) * h).sum())
    out = np.zeros(total, np.uint8)
    offs = np.zeros(p.nlevels + 1, np.int64)
    lib().vo_orb_pyramid(_p(gray), W, H, W, C.byref(p), _p(out), _p(offs))
    return [out[offs[l]:offs[l + 1]].reshape(h[l], w[l]) for l in range(p.nlevels)]


def orb_extract(gray, p, cap=20000):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    out = np.zeros(cap, KP_DTYPE)
    n = lib().vo_orb_extract(_p(gray), W, H, W, C.byref(p), _p(out), cap)
    assert n >= 0
    return out[:n].copy()


def orb_extract_describe(gray, p, cap=20000):
    """operator() with the descriptor call enabled: (key points, [n][32] uint8)"""
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    out = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = lib().vo_orb_extract_describe(_p(gray), W, H, W, C.byref(p), _p(out), cap, _p(desc))
    assert n >= 0
    return out[:n].copy(), desc[:n].copy()


def gauss7(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    lib().vo_gauss7_u8(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out), img.shape[1])
    return out


def describe_level(blurred, xs, ys, angles):
    blurred = np.ascontiguousarray(blurred, np.uint8)
    xs = np.ascontiguousarray(xs, np.float32); ys = np.ascontiguousarray(ys, np.float32)
    angles = np.ascontiguousarray(angles, np.float32)
    desc = np.zeros((len(xs), 32), np.uint8)
    lib().vo_orb_describe_level(_p(blurred), blurred.shape[1], blurred.shape[0], _p(xs), _p(ys), _p(angles), len(xs), _p(desc))
    return desc


def hamming_match(query, train):
    query = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
    train = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
    nq = len(query)
    bi = np.zeros(nq, np.int32); bd = np.zeros(nq, np.int32); sd = np.zeros(nq, np.int32)
    lib().vo_hamming_match(_p(query), nq, _p(train), len(train), _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


# ---------------------------------------------------------------- graph optimisation oracle
class LmRecord(C.Structure):
    _fields_ = [("chi2", C.c_double), ("lam", C.c_double), ("trials", C.c_int32), ("pad", C.c_int32)]


class LmStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("n_records", C.c_int32), ("total_trials", C.c_int32), ("pad", C.c_int32),
                ("rec", LmRecord * 320)]

    def records(self):
        return [(self.rec[i].chi2, self.rec[i].lam, self.rec[i].trials) for i in range(self.n_records)]


class BaProblem(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_points", C.c_int32), ("n_obs", C.c_int32), ("pad", C.c_int32),
                ("poses", C.c_void_p), ("rel_motion", C.c_void_p), ("points", C.c_void_p),
                ("obs_pose", C.c_void_p), ("obs_point", C.c_void_p), ("obs_xyz", C.c_void_p),
                ("max_iterations", C.c_int32), ("sigma2_cam", C.c_float), ("sigma2_3d", C.c_float),
                ("huber_cam", C.c_float), ("huber_3d", C.c_float), ("gain_threshold", C.c_float),
                ("fix_first", C.c_int32)]


def ba_partial(poses, rel_motion, points, obs_pose, obs_point, obs_xyz, **params):
    """Oracle sliding-window BA.  Arrays are copied; returns (poses, rel_motion, points, iterations, stats)."""
    poses = np.ascontiguousarray(poses, np.float32).copy()
    rel = np.ascontiguousarray(rel_motion, np.float32).copy()
    pts = np.ascontiguousarray(points, np.float32).copy()
    op = np.ascontiguousarray(obs_pose, np.int32)
    ol_ = np.ascontiguousarray(obs_point, np.int32)
    ox = np.ascontiguousarray(obs_xyz, np.float32)
    pr = BaProblem()
    lib().vo_ba_default_params(C.byref(pr))
    pr.n_poses, pr.n_points, pr.n_obs = len(poses), len(pts), len(op)
    pr.poses, pr.rel_motion, pr.points = _p(poses), _p(rel), _p(pts)
    pr.obs_pose, pr.obs_point, pr.obs_xyz = _p(op), _p(ol_), _p(ox)
    for k, v in params.items():
        setattr(pr, k, v)
    st = LmStats()
    its = lib().vo_ba_partial(C.byref(pr), C.byref(st))
    return poses, rel, pts, its, st


def edge_se3(Xi, Xj, Z):
    e = np.zeros(6); Ji = np.zeros((6, 6)); Jj = np.zeros((6, 6))
    lib().vo_edge_se3(_p(np.ascontiguousarray(Xi, np.float64)), _p(np.ascontiguousarray(Xj, np.float64)),
                      _p(np.ascontiguousarray(Z, np.float64)), _p(e), _p(Ji), _p(Jj))
    return e, Ji, Jj


def se3_oplus(X, u):
    out = np.zeros(12)
    lib().vo_se3_oplus(_p(np.ascontiguousarray(X, np.float64)), _p(np.ascontiguousarray(u, np.float64)), _p(out))
    return out


def edge_se3_pointxyz(X, p, z):
    e = np.zeros(3); Ji = np.zeros((3, 6)); Jj = np.zeros((3, 3))
    lib().vo_edge_se3_pointxyz(_p(np.ascontiguousarray(X, np.float64)), _p(np.ascontiguousarray(p, np.float64)),
                               _p(np.ascontiguousarray(z, np.float64)), _p(e), _p(Ji), _p(Jj))
    return e, Ji, Jj


# ---------------------------------------------------------------- per-frame stages
def depth_prep(depth, choose_data, factor, bf, mscale=1.0):
    d = np.ascontiguousarray(depth, np.float32).copy()
    lib().vo_depth_prep(_p(d), d.shape[1], d.shape[0], d.shape[1], int(choose_data), C.c_float(factor), C.c_float(bf),
                        C.c_float(mscale))
    return d


def frame_associate(kps, depth, flow, mask, th_depth_bg):
    kps = np.ascontiguousarray(kps)
    depth = np.ascontiguousarray(depth, np.float32); flow = np.ascontiguousarray(flow, np.float32)
    mask = np.ascontiguousarray(mask, np.int32)
    H, W = depth.shape
    cap = len(kps)
    idx = np.zeros(cap, np.int32); cor = np.zeros((cap, 2), np.float32); fl = np.zeros((cap, 2), np.float32)
    dep = np.zeros(cap, np.float32)
    n = lib().vo_frame_associate(_p(kps), len(kps), _p(depth), _p(flow), _p(mask), W, H, C.c_float(th_depth_bg),
                                 _p(idx), _p(cor), _p(fl), _p(dep), cap)
    return idx[:n].copy(), cor[:n].copy(), fl[:n].copy(), dep[:n].copy()


def frame_sample_objects(depth, flow, mask, th_depth_obj):
    depth = np.ascontiguousarray(depth, np.float32); flow = np.ascontiguousarray(flow, np.float32)
    mask = np.ascontiguousarray(mask, np.int32)
    H, W = depth.shape
    cap = ((H + 3) // 4) * ((W + 3) // 4)
    keys = np.zeros((cap, 2), np.float32); cor = np.zeros((cap, 2), np.float32); fl = np.zeros((cap, 2), np.float32)
    dep = np.zeros(cap, np.float32); lab = np.zeros(cap, np.int32)
    n = lib().vo_frame_sample_objects(_p(depth), _p(flow), _p(mask), W, H, C.c_float(th_depth_obj), _p(keys), _p(cor),
                                      _p(fl), _p(dep), _p(lab), cap)
    return keys[:n].copy(), cor[:n].copy(), fl[:n].copy(), dep[:n].copy(), lab[:n].copy()


def update_mask(sem_label, corres_xy, mask_last, flow_last, mask_cur):
    sem = np.ascontiguousarray(sem_label, np.int32); cor = np.ascontiguousarray(corres_xy, np.float32)
    ml = np.ascontiguousarray(mask_last, np.int32); fl = np.ascontiguousarray(flow_last, np.float32)
    mc = np.ascontiguousarray(mask_cur, np.int32).copy()
    H, W = ml.shape
    uniq = np.zeros(256, np.int32); rec = np.zeros(256, np.int32)
    k = lib().vo_update_mask(_p(sem), _p(cor), len(sem), _p(ml), _p(fl), _p(mc), W, H, _p(uniq), _p(rec), 256)
    return mc, uniq[:k].copy(), rec[:k].copy()


class PoseOptProblem(C.Structure):
    _fields_ = [("n", C.c_int32), ("pad", C.c_int32), ("obs_xy", C.c_void_p), ("flow_xy", C.c_void_p),
                ("depth", C.c_void_p), ("Tcw_init", C.c_float * 16), ("Tcw_last", C.c_float * 16),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("Tcw_out", C.c_float * 16), ("flow_out", C.c_void_p), ("inlier", C.c_void_p),
                ("info_flow", C.c_float), ("info_prior", C.c_float), ("rp_thres", C.c_float), ("chi2_th", C.c_float),
                ("rounds", C.c_int32), ("its", C.c_int32)]


def poseopt_flow2cam(obs_xy, flow_xy, depth, Tcw_init, Tcw_last, K, **params):
    """returns (Tcw 4x4 f32, refined flow [n,2], inlier mask [n], n_inliers, [LmStats]*rounds)"""
    obs = np.ascontiguousarray(obs_xy, np.float32); fl = np.ascontiguousarray(flow_xy, np.float32)
    dep = np.ascontiguousarray(depth, np.float32)
    n = len(obs)
    pr = PoseOptProblem()
    lib().vo_poseopt_default_params(C.byref(pr))
    pr.n = n
    pr.obs_xy, pr.flow_xy, pr.depth = _p(obs), _p(fl), _p(dep)
    pr.Tcw_init[:] = np.asarray(Tcw_init, np.float32).reshape(-1).tolist()
    pr.Tcw_last[:] = np.asarray(Tcw_last, np.float32).reshape(-1).tolist()
    pr.fx, pr.fy, pr.cx, pr.cy = [float(v) for v in K]
    fo = np.zeros((n, 2), np.float32); inl = np.zeros(n, np.int32)
    pr.flow_out, pr.inlier = _p(fo), _p(inl)
    for k, v in params.items():
        setattr(pr, k, v)
    stats = (LmStats * pr.rounds)()
    ninl = lib().vo_poseopt_flow2cam(C.byref(pr), stats)
    return np.array(pr.Tcw_out[:], np.float32).reshape(4, 4), fo, inl, ninl, list(stats)


class PnpProblem(C.Structure):
    _fields_ = [("n", C.c_int32), ("no_motion_model", C.c_int32), ("cur_xy", C.c_void_p), ("pts3d", C.c_void_p),
                ("valid", C.c_void_p), ("Tcw_motion", C.c_float * 16),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("iters", C.c_int32), ("reproj_err", C.c_float), ("confidence", C.c_float),
                ("Tcw_out", C.c_float * 16), ("inlier_ids", C.c_void_p),
                ("n_inliers", C.c_int32), ("winner", C.c_int32), ("ransac_inliers", C.c_int32), ("mm_inliers", C.c_int32)]


def fill_pnp(pr, cur_xy, pts3d, valid, Tcw_motion, K):
    cur = np.ascontiguousarray(cur_xy, np.float32); pts = np.ascontiguousarray(pts3d, np.float32)
    val = np.ascontiguousarray(valid, np.int32) if valid is not None else np.ones(len(cur), np.int32)
    ids = np.zeros(len(cur), np.int32)
    pr.n = len(cur)
    pr.cur_xy, pr.pts3d, pr.valid, pr.inlier_ids = _p(cur), _p(pts), _p(val), _p(ids)
    pr.Tcw_motion[:] = np.asarray(Tcw_motion, np.float32).reshape(-1).tolist()
    pr.fx, pr.fy, pr.cx, pr.cy = [float(v) for v in K]
    return cur, pts, val, ids


def init_model_cam(cur_xy, pts3d, valid, Tcw_motion, K, **params):
    """returns (Tcw 4x4 f32, inlier ids, winner, ransac_inliers, mm_inliers)"""
    pr = PnpProblem()
    lib().vo_pnp_default_params(C.byref(pr))
    keep = fill_pnp(pr, cur_xy, pts3d, valid, Tcw_motion, K)
    for k, v in params.items():
        setattr(pr, k, v)
    n = lib().vo_init_model_cam(C.byref(pr))
    return (np.array(pr.Tcw_out[:], np.float32).reshape(4, 4), keep[3][:n].copy(), pr.winner, pr.ransac_inliers,
            pr.mm_inliers)


# ---------------------------------------------------------------- full-sequence graph (FullBatchOptimization)
class FbaProblem(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_motions", C.c_int32), ("n_points", C.c_int32), ("n_obs", C.c_int32),
                ("n_e6", C.c_int32), ("n_tern", C.c_int32),
                ("se3", C.c_void_p), ("points", C.c_void_p),
                ("e6_i", C.c_void_p), ("e6_j", C.c_void_p), ("e6_kind", C.c_void_p), ("e6_meas", C.c_void_p),
                ("obs_se3", C.c_void_p), ("obs_point", C.c_void_p), ("obs_kind", C.c_void_p), ("obs_xyz", C.c_void_p),
                ("tern_p1", C.c_void_p), ("tern_p2", C.c_void_p), ("tern_h", C.c_void_p),
                ("max_iterations", C.c_int32),
                ("sigma2_cam", C.c_float), ("sigma2_3d_sta", C.c_float), ("sigma2_3d_dyn", C.c_float),
                ("sigma2_obj", C.c_float), ("sigma2_smooth", C.c_float),
                ("huber_cam", C.c_float), ("huber_obj", C.c_float), ("huber_3d", C.c_float),
                ("gain_threshold", C.c_float), ("prior_info", C.c_float)]


FBA_KEYS = ("se3", "points", "e6_i", "e6_j", "e6_kind", "e6_meas", "obs_se3", "obs_point", "obs_kind", "obs_xyz",
            "tern_p1", "tern_p2", "tern_h")


def fill_fba(pr, g, n_poses):
    """g: dict of numpy arrays (FBA_KEYS); the arrays se3 / points are updated in place by the solver"""
    f32 = ("se3", "points", "e6_meas", "obs_xyz")
    keep = {k: np.ascontiguousarray(g[k], np.float32 if k in f32 else np.int32) for k in FBA_KEYS}
    pr.n_poses = n_poses
    pr.n_motions = keep["se3"].reshape(-1, 16).shape[0] - n_poses
    pr.n_points = keep["points"].reshape(-1, 3).shape[0]
    pr.n_obs, pr.n_e6, pr.n_tern = len(keep["obs_se3"]), len(keep["e6_i"]), len(keep["tern_p1"])
    for k in FBA_KEYS:
        setattr(pr, k, _p(keep[k]).value if keep[k].size else None)
    return keep


def ba_full(g, n_poses, **params):
    """Optimizer::FullBatchOptimization on a flat graph; returns (se3 [n,4,4], points, iterations, LmStats)"""
    pr = FbaProblem()
    lib().vo_fba_default_params(C.byref(pr))
    keep = fill_fba(pr, {k: np.array(g[k], copy=True) for k in FBA_KEYS}, n_poses)
    for k, v in params.items():
        setattr(pr, k, v)
    st = LmStats()
    its = lib().vo_ba_full(C.byref(pr), C.byref(st))
    return keep["se3"].reshape(-1, 4, 4), keep["points"].reshape(-1, 3), its, st


def edge_landmark_motion(H, p1, p2):
    H = np.ascontiguousarray(H, np.float64); p1 = np.ascontiguousarray(p1, np.float64); p2 = np.ascontiguousarray(p2, np.float64)
    e = np.zeros(3); J2 = np.zeros((3, 3)); JH = np.zeros((3, 6))
    lib().vo_edge_landmark_motion(_p(H), _p(p1), _p(p2), _p(e), _p(J2), _p(JH))
    return e, J2, JH


def edge_se3_prior(X, Z):
    X = np.ascontiguousarray(X, np.float64); Z = np.ascontiguousarray(Z, np.float64)
    e = np.zeros(6); J = np.zeros((6, 6))
    lib().vo_edge_se3_prior(_p(X), _p(Z), _p(e), _p(J))
    return e, J


# ---------------------------------------------------------------- tracking pipeline oracle
class TrackConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("bf", C.c_float), ("choose_data", C.c_int32), ("depth_map_factor", C.c_float),
                ("th_depth_bg", C.c_float), ("th_depth_obj", C.c_float), ("max_track_bg", C.c_int32),
                ("window_size", C.c_int32), ("orb", OrbParams), ("rebuild_tracklets", C.c_int32),
                ("max_track_obj", C.c_int32), ("sf_mg_thres", C.c_float), ("sf_ds_thres", C.c_float), ("b_joint", C.c_int32)]


class TrackStats(C.Structure):
    _fields_ = [("ms_orb", C.c_double), ("ms_assoc", C.c_double), ("ms_init", C.c_double), ("ms_poseopt", C.c_double),
                ("ms_renew", C.c_double), ("ms_ba", C.c_double),
                ("n_keypoints", C.c_int32), ("n_matches", C.c_int32), ("n_init_inliers", C.c_int32),
                ("init_winner", C.c_int32), ("n_pose_inliers", C.c_int32), ("n_static", C.c_int32),
                ("ba_iterations", C.c_int32), ("ba_trials", C.c_int32), ("ba_points", C.c_int32), ("ba_obs", C.c_int32),
                ("n_dyn_features", C.c_int32), ("n_objects", C.c_int32), ("n_objects_ok", C.c_int32),
                ("n_masks_recovered", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def track_config(cam, nfeatures=2500, window=20, max_track_bg=1000, rebuild=1, choose_data=2, depth_map_factor=256.0,
                 th_depth_bg=5000.0, th_depth_obj=25.0, max_track_obj=500, sf_mg_thres=0.12, sf_ds_thres=0.3, b_joint=1):
    c = TrackConfig()
    c.width, c.height = cam["width"], cam["height"]
    c.fx, c.fy, c.cx, c.cy, c.bf = cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["bf"]
    c.choose_data, c.depth_map_factor, c.th_depth_bg, c.th_depth_obj = choose_data, depth_map_factor, th_depth_bg, th_depth_obj
    c.max_track_bg, c.window_size = max_track_bg, window
    c.orb = default_orb_params(nfeatures)
    c.rebuild_tracklets = rebuild
    c.max_track_obj, c.sf_mg_thres, c.sf_ds_thres = max_track_obj, sf_mg_thres, sf_ds_thres
    c.b_joint = b_joint
    return c


def renew_static(cfg, TM_sta, stat_keys, kps, depth, flow, mask, Tcw):
    """static half of Tracking::RenewFrameInfo of the oracle on caller-supplied state: keys, corres, flow, inlier ids, depth, 3-D points"""
    L = lib()
    L.vo_renew_static.argtypes = [C.POINTER(TrackConfig), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 10
    tm = np.ascontiguousarray(TM_sta, np.int32); sk = np.ascontiguousarray(stat_keys, np.float32); kp = np.ascontiguousarray(kps, KP_DTYPE)
    depth = np.ascontiguousarray(depth, np.float32); flow = np.ascontiguousarray(flow, np.float32); mask = np.ascontiguousarray(mask, np.int32)
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16)
    cap = cfg.max_track_bg + 2
    keys = np.zeros((cap, 2), np.float32); cor = np.zeros((cap, 2), np.float32); fl = np.zeros((cap, 2), np.float32)
    inl = np.zeros(cap, np.int32); dep = np.zeros(cap, np.float32); p3 = np.zeros((cap, 3), np.float32)
    n = L.vo_renew_static(C.byref(cfg), _p(tm), len(tm), _p(sk), len(sk), _p(kp), len(kp), _p(depth), _p(flow), _p(mask), _p(T), _p(keys),
                          _p(cor), _p(fl), _p(inl), _p(dep), _p(p3))
    assert n <= cap
    return keys[:n], cor[:n], fl[:n], inl[:n], dep[:n], p3[:n]


def renew_objects(cfg, obj_keys, obj_label, inlier_sets, obj_stat, sem_pos, mod_label, tmp, depth, flow, mask, Tcw, cap=1 << 16):
    """object half of Tracking::RenewFrameInfo of the oracle on caller-supplied state; tmp = dict(keys, depth, sem, flow, corres) of the
    frame's object samples.  Returns keys, depth, corres, flow, semantic label, inlier id, object label, 3-D points."""
    L = lib()
    L.vo_renew_objects.argtypes = ([C.POINTER(TrackConfig), C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int] +
                                   [C.c_void_p] * 9 + [C.c_int] + [C.c_void_p] * 8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    ok, ol_ = f32(obj_keys), i32(obj_label)
    ilen = i32([len(s) for s in inlier_sets]); iids = i32(np.concatenate([np.asarray(s, np.int32) for s in inlier_sets]) if len(inlier_sets) else [])
    st, sp, ml = i32(obj_stat), i32(sem_pos), i32(mod_label)
    tk, td, ts, tf, tc = f32(tmp["keys"]), f32(tmp["depth"]), i32(tmp["sem"]), f32(tmp["flow"]), f32(tmp["corres"])
    depth, flow, mask, T = f32(depth), f32(flow), i32(mask), f32(Tcw).reshape(16)
    keys = np.zeros((cap, 2), np.float32); dep = np.zeros(cap, np.float32); cor = np.zeros((cap, 2), np.float32); fl = np.zeros((cap, 2), np.float32)
    sem = np.zeros(cap, np.int32); inl = np.zeros(cap, np.int32); lab = np.zeros(cap, np.int32); p3 = np.zeros((cap, 3), np.float32)
    n = L.vo_renew_objects(C.byref(cfg), len(ol_), _p(ok), _p(ol_), len(ilen), _p(ilen), _p(iids), _p(st), _p(sp), _p(ml), len(ts), _p(tk),
                           _p(td), _p(ts), _p(tf), _p(tc), _p(depth), _p(flow), _p(mask), _p(T), cap, _p(keys), _p(dep), _p(cor), _p(fl),
                           _p(sem), _p(inl), _p(lab), _p(p3))
    assert 0 <= n <= cap
    return keys[:n], dep[:n], cor[:n], fl[:n], sem[:n], inl[:n], lab[:n], p3[:n]


def dyn_obj_tracking(cfg, sem, lab, key_xy, depth, flow3, last_sem, last_sem_pos, last_stat, last_mod, f_id, max_id):
    """Tracking::DynObjTracking of the oracle on caller-supplied frame state: (labels after, max_id after, nModLabel, nSemPosition,
    list of index arrays)"""
    L = lib()
    L.vo_dyn_obj_tracking.argtypes = [C.POINTER(TrackConfig), C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 5
    n = len(sem)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    sem, lab, last_sem = i32(sem), i32(lab).copy(), i32(last_sem)
    key_xy, depth, flow3 = f32(key_xy), f32(depth), f32(flow3)
    lsp, lst, lmd = i32(last_sem_pos), i32(last_stat), i32(last_mod)
    mx = np.array([max_id], np.int32)
    out_mod = np.zeros(n + 1, np.int32); out_sem = np.zeros(n + 1, np.int32); out_len = np.zeros(n + 1, np.int32); out_ids = np.zeros(n + 1, np.int32)
    no = L.vo_dyn_obj_tracking(C.byref(cfg), n, _p(sem), _p(lab), _p(key_xy), _p(depth), _p(flow3), _p(last_sem), len(lsp), _p(lsp), _p(lst),
                               _p(lmd), f_id, _p(mx), _p(out_mod), _p(out_sem), _p(out_len), _p(out_ids))
    ids, at = [], 0
    for o in range(no):
        ids.append(out_ids[at:at + out_len[o]].copy()); at += out_len[o]
    return lab, int(mx[0]), out_mod[:no].copy(), out_sem[:no].copy(), ids


class Metric(C.Structure):
    """vo_metric / vido_metric"""
    _fields_ = [("cam_t", C.c_float), ("cam_r", C.c_float), ("obj_t", C.c_float), ("obj_r", C.c_float), ("n_cam", C.c_int32),
                ("n_obj", C.c_int32)]


class ImuState(C.Structure):
    """vo_imu_state / vido_imu_state"""
    _fields_ = [("initialized", C.c_int32), ("status", C.c_int32), ("init_frame", C.c_int32), ("n_refinements", C.c_int32),
                ("n_reintegrated", C.c_int32), ("lm_iterations", C.c_int32), ("lm_trials", C.c_int32), ("t_init", C.c_float),
                ("scale", C.c_double), ("Rwg", C.c_double * 9), ("bg", C.c_double * 3), ("ba", C.c_double * 3)]


class OracleTracker:
    def __init__(self, cfg):
        L = lib()
        L.vo_tracker_create.restype = C.c_void_p
        L.vo_tracker_create.argtypes = [C.POINTER(TrackConfig)]
        L.vo_tracker_destroy.argtypes = [C.c_void_p]
        L.vo_tracker_track.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.POINTER(TrackStats)]
        L.vo_tracker_get_map_poses.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.vo_tracker_num_frames.argtypes = [C.c_void_p]
        L.vo_tracker_get_static.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int]
        L.vo_tracker_get_dynamic.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int]
        L.vo_tracker_get_objects.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int]
        L.vo_tracker_get_dyn_tracks.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int]
        L.vo_tracker_full_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vo_tracker_get_map_poses_rf.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.vo_tracker_get_objects_rf.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.vo_tracker_export_full_graph.argtypes = [C.c_void_p] + [C.c_void_p] * 14
        L.vo_tracker_set_imu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vo_tracker_grab_imu.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.vo_tracker_set_timestamp.argtypes = [C.c_void_p, C.c_double]
        L.vo_tracker_set_timestamp.restype = None
        L.vo_tracker_get_imu_state.argtypes = [C.c_void_p, C.c_void_p]
        L.vo_tracker_get_imu_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.vo_tracker_apply_scaled_rotation.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        self.cfg = cfg
        self.h = L.vo_tracker_create(C.byref(cfg))

    # ---- recorded Tracking::DynObjTracking calls (inputs and outputs), for tests/test_dynobj_independent.py
    def dyn_log_enable(self, on=True):
        lib().vo_tracker_dyn_log_enable.argtypes = [C.c_void_p, C.c_int]
        lib().vo_tracker_dyn_log_enable.restype = None
        lib().vo_tracker_dyn_log_enable(self.h, 1 if on else 0)

    def dyn_log(self):
        L = lib()
        L.vo_tracker_dyn_log_count.argtypes = [C.c_void_p]
        L.vo_tracker_dyn_log_sizes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.vo_tracker_dyn_log_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 14
        out = []
        for k in range(L.vo_tracker_dyn_log_count(self.h)):
            sz = np.zeros(8, np.int32)
            L.vo_tracker_dyn_log_sizes(self.h, k, _p(sz))
            n, nl, nlo, no, nid = [int(v) for v in sz[:5]]
            assert nl == n
            r = dict(f_id=int(sz[5]), max_id_before=int(sz[6]), max_id_after=int(sz[7]),
                     sem=np.zeros(n, np.int32), lab_before=np.zeros(n, np.int32), lab_after=np.zeros(n, np.int32),
                     last_sem=np.zeros(n, np.int32), key_xy=np.zeros((n, 2), np.float32), depth=np.zeros(n, np.float32),
                     flow3=np.zeros((n, 3), np.float32), last_sem_pos=np.zeros(nlo, np.int32), last_stat=np.zeros(nlo, np.int32),
                     last_mod=np.zeros(nlo, np.int32), out_mod=np.zeros(no, np.int32), out_sem_pos=np.zeros(no, np.int32),
                     out_len=np.zeros(no, np.int32), out_ids=np.zeros(nid, np.int32))
            L.vo_tracker_dyn_log_get(self.h, k, *[_p(r[key]) for key in ("sem", "lab_before", "lab_after", "last_sem", "key_xy", "depth", "flow3",
                                                                        "last_sem_pos", "last_stat", "last_mod", "out_mod", "out_sem_pos",
                                                                        "out_len", "out_ids")])
            out.append(r)
        return out

    def imu_init_log(self):
        """recorded gravity-direction / velocity initialisations of Tracking::InitializeIMU (needs dyn_log_enable before tracking)"""
        L = lib()
        L.vo_tracker_imu_init_log.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        out = []
        for k in range(L.vo_tracker_imu_init_log(self.h, -1, *([None] * 7))):
            n = L.vo_tracker_imu_init_log(self.h, k, *([None] * 7))
            r = dict(Tcw=np.zeros((n, 4, 4), np.float32), has_pre=np.zeros(n, np.int32), dV=np.zeros((n, 3), np.float32),
                     dT=np.zeros(n, np.float32), Tbc=np.zeros((4, 4), np.float32), Rwg=np.zeros((3, 3), np.float32),
                     vel_out=np.zeros((n, 3), np.float32))
            L.vo_tracker_imu_init_log(self.h, k, *[_p(r[key]) for key in ("Tcw", "has_pre", "dV", "dT", "Tbc", "Rwg", "vel_out")])
            out.append(r)
        return out

    def imu_update_log(self, cap=64):
        """recorded Tracking::UpdateFrameIMU calls: dicts of Rwb1, twb1, Vwb1, dR, dV, dP, t12 (inputs) and Rwb, twb, Vwb (outputs)"""
        L = lib()
        L.vo_tracker_imu_update_log.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        buf = np.zeros((cap, 46), np.float32)
        n = L.vo_tracker_imu_update_log(self.h, _p(buf), cap)
        out = []
        for r in buf[:n]:
            out.append(dict(Rwb1=r[0:9].reshape(3, 3), twb1=r[9:12], Vwb1=r[12:15], dR=r[15:24].reshape(3, 3), dV=r[24:27], dP=r[27:30],
                            t12=r[30], Rwb=r[31:40].reshape(3, 3), twb=r[40:43], Vwb=r[43:46]))
        return out

    # ---- VIO mode (sensor = IMU_RGBD)
    def set_imu(self, Tbc, noise):
        T = np.ascontiguousarray(Tbc, np.float32).reshape(16); nz = np.ascontiguousarray(noise, np.float32)
        lib().vo_tracker_set_imu(self.h, _p(T), _p(nz))

    def grab_imu(self, samples):
        s = np.ascontiguousarray(samples, IMU_SAMPLE)
        lib().vo_tracker_grab_imu(self.h, _p(s), len(s))

    def imu_state(self):
        st = ImuState()
        lib().vo_tracker_get_imu_state(self.h, C.byref(st))
        return st

    def imu_frames(self):
        n = lib().vo_tracker_get_imu_frames(self.h, None, None, None, 0)
        T = np.zeros((n, 16), np.float32); v = np.zeros((n, 3), np.float32); b = np.zeros((n, 6), np.float32)
        lib().vo_tracker_get_imu_frames(self.h, _p(T), _p(v), _p(b), n)
        return T.reshape(n, 4, 4), v, b

    def apply_scaled_rotation(self, R, s):
        Rm = np.ascontiguousarray(R, np.float32).reshape(9)
        lib().vo_tracker_apply_scaled_rotation(self.h, _p(Rm), float(s))

    def metric_error(self, cam_gt, obj_pose_pre=None, obj_motion_gt=None, refined=False):
        """Tracking::GetMetricError; returns (Metric, per-item (t, r) array)"""
        L = lib()
        L.vo_tracker_metric_error.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                              C.POINTER(Metric), C.c_void_p]
        g = np.ascontiguousarray(cam_gt, np.float32).reshape(-1, 16)
        no = 0 if obj_pose_pre is None else len(obj_pose_pre)
        pp = np.ascontiguousarray(obj_pose_pre, np.float32).reshape(-1, 16) if no else None
        mg = np.ascontiguousarray(obj_motion_gt, np.float32).reshape(-1, 16) if no else None
        m = Metric()
        per = np.zeros((max(len(g) - 1, 0) + no, 2), np.float32)
        rc = L.vo_tracker_metric_error(self.h, _p(g), len(g), int(refined), _p(pp) if no else None, _p(mg) if no else None, no,
                                       C.byref(m), _p(per))
        assert rc == 0, rc
        return m, per

    def track(self, gray, depth_in, flow, mask, timestamp=None):
        """depth_in is copied (the oracle pre-scales its copy in place).  Returns (Tcw 4x4, stats dict, rc)."""
        if timestamp is not None:
            lib().vo_tracker_set_timestamp(self.h, float(timestamp))
        g = np.ascontiguousarray(gray, np.uint8); d = np.ascontiguousarray(depth_in, np.float32).copy()
        f = np.ascontiguousarray(flow, np.float32); m = np.ascontiguousarray(mask, np.int32)
        T = np.zeros(16, np.float32)
        st = TrackStats()
        rc = lib().vo_tracker_track(self.h, _p(g), _p(d), _p(f), _p(m), _p(T), C.byref(st))
        return T.reshape(4, 4), st.as_dict(), rc

    def map_poses(self):
        n = lib().vo_tracker_num_frames(self.h)
        P = np.zeros((n, 16), np.float32)
        lib().vo_tracker_get_map_poses(self.h, _p(P), n)
        return P.reshape(n, 4, 4)

    def static_features(self, frame, cap=4096):
        xy = np.zeros((cap, 2), np.float32); dep = np.zeros(cap, np.float32); p3 = np.zeros((cap, 3), np.float32)
        asso = np.zeros(cap, np.int32)
        n = lib().vo_tracker_get_static(self.h, frame, _p(xy), _p(dep), _p(p3), _p(asso), cap)
        return xy[:n].copy(), dep[:n].copy(), p3[:n].copy(), asso[:n].copy()

    def dynamic_features(self, frame, cap=32768):
        """Map::vpFeatDyn / vfDepDyn / vp3DPointDyn / vnAssoDyn / vnFeatLabel of one frame"""
        xy = np.zeros((cap, 2), np.float32); dep = np.zeros(cap, np.float32); p3 = np.zeros((cap, 3), np.float32)
        asso = np.zeros(cap, np.int32); lab = np.zeros(cap, np.int32)
        n = lib().vo_tracker_get_dynamic(self.h, frame, _p(xy), _p(dep), _p(p3), _p(asso), _p(lab), cap)
        return xy[:n].copy(), dep[:n].copy(), p3[:n].copy(), asso[:n].copy(), lab[:n].copy()

    def objects(self, frame, cap=64):
        """(tracking label, semantic label, motion 4x4, centre) of the objects with an estimated motion in `frame` (>= 1)"""
        lab = np.zeros(cap, np.int32); sem = np.zeros(cap, np.int32); mot = np.zeros((cap, 16), np.float32)
        cen = np.zeros((cap, 3), np.float32)
        n = max(lib().vo_tracker_get_objects(self.h, frame, _p(lab), _p(sem), _p(mot), _p(cen), cap), 0)
        return lab[:n].copy(), sem[:n].copy(), mot[:n].reshape(n, 4, 4).copy(), cen[:n].copy()

    def dyn_tracks(self, cap=1 << 20):
        ln = np.zeros(cap, np.int32); oid = np.zeros(cap, np.int32); ff = np.zeros(cap, np.int32); fj = np.zeros(cap, np.int32)
        n = lib().vo_tracker_get_dyn_tracks(self.h, _p(ln), _p(oid), _p(ff), _p(fj), cap)
        return ln[:n].copy(), oid[:n].copy(), ff[:n].copy(), fj[:n].copy()

    def full_batch(self):
        """Optimizer::FullBatchOptimization on the map; returns (iterations, LmStats, sizes[6])"""
        st = LmStats(); sizes = np.zeros(6, np.int32)
        its = lib().vo_tracker_full_batch(self.h, C.byref(st), _p(sizes))
        return its, st, sizes

    def map_poses_rf(self):
        n = lib().vo_tracker_num_frames(self.h)
        P = np.zeros((n, 16), np.float32)
        lib().vo_tracker_get_map_poses_rf(self.h, _p(P), n)
        return P.reshape(n, 4, 4)

    def objects_rf(self, frame, cap=64):
        mot = np.zeros((cap, 16), np.float32)
        n = max(lib().vo_tracker_get_objects_rf(self.h, frame, _p(mot), cap), 0)
        return mot[:n].reshape(n, 4, 4).copy()

    def export_full_graph(self):
        """the flat FullBatch graph of the current map: (dict of arrays keyed by FBA_KEYS, n_poses)"""
        sizes = np.zeros(6, np.int32)
        lib().vo_tracker_export_full_graph(self.h, _p(sizes), *([None] * 13))
        npo, nmo, npt, nob, ne6, nte = [int(v) for v in sizes]
        g = dict(se3=np.zeros((npo + nmo, 16), np.float32), points=np.zeros((npt, 3), np.float32),
                 e6_i=np.zeros(ne6, np.int32), e6_j=np.zeros(ne6, np.int32), e6_kind=np.zeros(ne6, np.int32),
                 e6_meas=np.zeros((ne6, 16), np.float32), obs_se3=np.zeros(nob, np.int32), obs_point=np.zeros(nob, np.int32),
                 obs_kind=np.zeros(nob, np.int32), obs_xyz=np.zeros((nob, 3), np.float32), tern_p1=np.zeros(nte, np.int32),
                 tern_p2=np.zeros(nte, np.int32), tern_h=np.zeros(nte, np.int32))
        lib().vo_tracker_export_full_graph(self.h, _p(sizes), *[_p(g[k]) if g[k].size else None for k in FBA_KEYS])
        return g, npo

    def close(self):
        if self.h:
            lib().vo_tracker_destroy(self.h)
            self.h = None


# ---------------------------------------------------------------- IMU preintegration oracle
IMU_SAMPLE = np.dtype([("t", "<f8"), ("ax", "<f4"), ("ay", "<f4"), ("az", "<f4"), ("wx", "<f4"), ("wy", "<f4"), ("wz", "<f4")])
IMU_PREINT = np.dtype([("dT", "<f4"), ("dR", "<f4", 9), ("dV", "<f4", 3), ("dP", "<f4", 3), ("JRg", "<f4", 9),
                       ("JVg", "<f4", 9), ("JVa", "<f4", 9), ("JPg", "<f4", 9), ("JPa", "<f4", 9), ("C", "<f4", 225),
                       ("avgA", "<f4", 3), ("avgW", "<f4", 3), ("n_steps", "<i4"), ("n_consumed", "<i4")])


def imu_preintegrate(samples, t_prev, t_cur, bias, noise):
    s = np.ascontiguousarray(samples, IMU_SAMPLE)
    b = np.ascontiguousarray(bias, np.float32); nz = np.ascontiguousarray(noise, np.float32)
    out = np.zeros(1, IMU_PREINT)
    L = lib()
    L.vo_imu_preintegrate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.vo_imu_preintegrate(_p(s), len(s), float(t_prev), float(t_cur), _p(b), _p(nz), _p(out))
    return out[0]


# ---------------------------------------------------------------- inertial-only optimisation (VIO initialisation)
class InertialProblem(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("its", C.c_int32), ("Rwb", C.c_void_p), ("twb", C.c_void_p), ("velocity", C.c_void_p),
                ("preint", C.c_void_p), ("bias_lin", C.c_void_p), ("Rwg", C.c_double * 9), ("scale", C.c_double),
                ("bg", C.c_double * 3), ("ba", C.c_double * 3), ("prior_g", C.c_float), ("prior_a", C.c_float),
                ("mode", C.c_int32)]


def fill_inertial(pr, Rwb, twb, vel, preint, bias_lin, Rwg, scale, bg, ba):
    keep = [np.ascontiguousarray(Rwb, np.float32).reshape(-1, 9), np.ascontiguousarray(twb, np.float32).reshape(-1, 3),
            np.array(vel, np.float32).reshape(-1, 3).copy(), np.ascontiguousarray(preint, IMU_PREINT),
            np.ascontiguousarray(bias_lin, np.float32).reshape(-1, 6)]
    pr.n_frames = keep[0].shape[0]
    pr.Rwb, pr.twb, pr.velocity, pr.preint, pr.bias_lin = [_p(a).value for a in keep]
    pr.Rwg[:] = [float(v) for v in np.asarray(Rwg, np.float64).reshape(9)]
    pr.scale = float(scale)
    pr.bg[:] = [float(v) for v in bg]
    pr.ba[:] = [float(v) for v in ba]
    return keep


def inertial_optimization(Rwb, twb, vel, preint, bias_lin, Rwg, scale=1.0, bg=(0, 0, 0), ba=(0, 0, 0), **params):
    """Optimizer::InertialOptimization; returns dict(velocity, Rwg, scale, bg, ba, iterations, stats)"""
    pr = InertialProblem()
    lib().vo_inertial_default_params(C.byref(pr))
    keep = fill_inertial(pr, Rwb, twb, vel, preint, bias_lin, Rwg, scale, bg, ba)
    for k, v in params.items():
        setattr(pr, k, v)
    st = LmStats()
    its = lib().vo_inertial_optimization(C.byref(pr), C.byref(st))
    return dict(velocity=keep[2], Rwg=np.array(pr.Rwg[:]).reshape(3, 3), scale=pr.scale, bg=np.array(pr.bg[:]), ba=np.array(pr.ba[:]),
                iterations=its, stats=st)


def inertial_edge_information(C15):
    Cm = np.ascontiguousarray(C15, np.float32).reshape(15, 15)
    out = np.zeros((9, 9))
    lib().vo_inertial_edge_information(_p(Cm), _p(out))
    return out


# ---------------------------------------------------------------- reprojection-only optimisers (bJoint == false)
class ProjOptProblem(C.Structure):
    _fields_ = [("n", C.c_int32), ("kind", C.c_int32), ("obs_xy", C.c_void_p), ("pts3d", C.c_void_p),
                ("T_init", C.c_float * 16), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("P", C.c_double * 12), ("rp_thres", C.c_float), ("its", C.c_int32), ("T_out", C.c_float * 16),
                ("inlier", C.c_void_p), ("n_inliers", C.c_int32)]


def fill_projopt(pr, kind, obs_xy, pts3d, T_init, K=None, P=None):
    keep = [np.ascontiguousarray(obs_xy, np.float32).reshape(-1, 2), np.ascontiguousarray(pts3d, np.float32).reshape(-1, 3)]
    keep.append(np.zeros(len(keep[0]), np.int32))
    pr.n = len(keep[0])
    pr.obs_xy, pr.pts3d, pr.inlier = _p(keep[0]).value, _p(keep[1]).value, _p(keep[2]).value
    pr.T_init[:] = [float(v) for v in np.asarray(T_init, np.float32).reshape(16)]
    if K is not None:
        pr.fx, pr.fy, pr.cx, pr.cy = K
    if P is not None:
        pr.P[:] = [float(v) for v in np.asarray(P, np.float64).reshape(12)]
    return keep


def pose_opt_proj(kind, obs_xy, pts3d, T_init, K=None, P=None, **params):
    """kind 0: Optimizer::PoseOptimizationNew, kind 1: PoseOptimizationObjMot.  Returns (T 4x4, inlier flags, LmStats)"""
    pr = ProjOptProblem()
    lib().vo_projopt_default_params(C.byref(pr), kind)
    keep = fill_projopt(pr, kind, obs_xy, pts3d, T_init, K, P)
    for k, v in params.items():
        setattr(pr, k, v)
    st = LmStats()
    lib().vo_pose_opt_proj(C.byref(pr), C.byref(st))
    return np.array(pr.T_out[:], np.float32).reshape(4, 4), keep[2].copy(), st
