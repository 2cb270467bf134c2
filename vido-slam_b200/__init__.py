"""ctypes binding of libvido_b200.so -- thin: every call goes straight to the C-ABI in include/vido_b200.h.

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but creating a
context raises when the CUDA library or an sm_100 device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VIDO_LIB_PATH", os.path.join(_HERE, "libvido_b200.so"))  # override: debug builds only

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])


class VidoConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("choose_data", C.c_int32), ("depth_map_factor", C.c_float),
                ("th_depth_bg", C.c_float), ("th_depth_obj", C.c_float),
                ("max_track_bg", C.c_int32), ("max_track_obj", C.c_int32), ("window_size", C.c_int32),
                ("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("rgb", C.c_int32),
                ("max_batch", C.c_int32), ("device", C.c_int32), ("sf_mg_thres", C.c_float), ("sf_ds_thres", C.c_float),
                ("b_joint", C.c_int32)]


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


class PoseOptProblem(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_inliers", C.c_int32), ("obs_xy", C.c_void_p), ("flow_xy", C.c_void_p),
                ("depth", C.c_void_p), ("Tcw_init", C.c_float * 16), ("Tcw_last", C.c_float * 16),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("Tcw_out", C.c_float * 16), ("flow_out", C.c_void_p), ("inlier", C.c_void_p),
                ("info_flow", C.c_float), ("info_prior", C.c_float), ("rp_thres", C.c_float), ("chi2_th", C.c_float),
                ("rounds", C.c_int32), ("its", C.c_int32)]


class PnpProblem(C.Structure):
    _fields_ = [("n", C.c_int32), ("no_motion_model", C.c_int32), ("cur_xy", C.c_void_p), ("pts3d", C.c_void_p),
                ("valid", C.c_void_p), ("Tcw_motion", C.c_float * 16),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("iters", C.c_int32), ("reproj_err", C.c_float), ("confidence", C.c_float),
                ("Tcw_out", C.c_float * 16), ("inlier_ids", C.c_void_p),
                ("n_inliers", C.c_int32), ("winner", C.c_int32), ("ransac_inliers", C.c_int32), ("mm_inliers", C.c_int32)]


class FbaProblem(C.Structure):
    """vido_fba_problem (include/vido_b200.h): flat graph of Optimizer::FullBatchOptimization"""
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
                ("gain_threshold", C.c_float), ("prior_info", C.c_float), ("solver", C.c_int32)]


FBA_KEYS = ("se3", "points", "e6_i", "e6_j", "e6_kind", "e6_meas", "obs_se3", "obs_point", "obs_kind", "obs_xyz",
            "tern_p1", "tern_p2", "tern_h")
FBA_F32 = ("se3", "points", "e6_meas", "obs_xyz")
FBA_WIDTH = dict(se3=16, points=3, e6_meas=16, obs_xyz=3)


class ProjOptProblem(C.Structure):
    """vido_projopt_problem (include/vido_b200.h)"""
    _fields_ = [("n", C.c_int32), ("kind", C.c_int32), ("obs_xy", C.c_void_p), ("pts3d", C.c_void_p),
                ("T_init", C.c_float * 16), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("P", C.c_double * 12), ("rp_thres", C.c_float), ("its", C.c_int32), ("T_out", C.c_float * 16),
                ("inlier", C.c_void_p), ("n_inliers", C.c_int32)]


class InertialProblem(C.Structure):
    """vido_inertial_problem (include/vido_b200.h)"""
    _fields_ = [("n_frames", C.c_int32), ("its", C.c_int32), ("Rwb", C.c_void_p), ("twb", C.c_void_p), ("velocity", C.c_void_p),
                ("preint", C.c_void_p), ("bias_lin", C.c_void_p), ("Rwg", C.c_double * 9), ("scale", C.c_double),
                ("bg", C.c_double * 3), ("ba", C.c_double * 3), ("prior_g", C.c_float), ("prior_a", C.c_float),
                ("mode", C.c_int32)]


class Metric(C.Structure):
    """vido_metric (include/vido_b200.h)"""
    _fields_ = [("cam_t", C.c_float), ("cam_r", C.c_float), ("obj_t", C.c_float), ("obj_r", C.c_float), ("n_cam", C.c_int32),
                ("n_obj", C.c_int32)]


class ImuState(C.Structure):
    """vido_imu_state (include/vido_b200.h)"""
    _fields_ = [("initialized", C.c_int32), ("status", C.c_int32), ("init_frame", C.c_int32), ("n_refinements", C.c_int32),
                ("n_reintegrated", C.c_int32), ("lm_iterations", C.c_int32), ("lm_trials", C.c_int32), ("t_init", C.c_float),
                ("scale", C.c_double), ("Rwg", C.c_double * 9), ("bg", C.c_double * 3), ("ba", C.c_double * 3)]


class FrameInputs(C.Structure):
    _fields_ = [("image", C.c_void_p), ("channels", C.c_int32), ("on_device", C.c_int32), ("depth", C.c_void_p),
                ("flow", C.c_void_p), ("mask", C.c_void_p), ("write_back_depth", C.c_int32), ("pad", C.c_int32),
                ("timestamp", C.c_double)]


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


IMU_SAMPLE = np.dtype([("t", "<f8"), ("ax", "<f4"), ("ay", "<f4"), ("az", "<f4"), ("wx", "<f4"), ("wy", "<f4"), ("wz", "<f4")])
IMU_PREINT = np.dtype([("dT", "<f4"), ("dR", "<f4", 9), ("dV", "<f4", 3), ("dP", "<f4", 3), ("JRg", "<f4", 9),
                       ("JVg", "<f4", 9), ("JVa", "<f4", 9), ("JPg", "<f4", 9), ("JPa", "<f4", 9), ("C", "<f4", 225),
                       ("avgA", "<f4", 3), ("avgW", "<f4", 3), ("n_steps", "<i4"), ("n_consumed", "<i4")])


class VidoError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VidoError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a)")
    lib = C.CDLL(LIB_PATH)
    vp, i32p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)
    lib.vido_version.restype = C.c_int
    lib.vido_default_config.argtypes = [C.POINTER(VidoConfig)]
    lib.vido_create.restype = vp
    lib.vido_create.argtypes = [C.POINTER(VidoConfig)]
    lib.vido_destroy.argtypes = [vp]
    lib.vido_last_error.restype = C.c_char_p
    lib.vido_last_error.argtypes = [vp]
    lib.vido_kernel_launches.restype = C.c_int64
    lib.vido_kernel_launches.argtypes = [vp]
    lib.vido_stream.restype = vp
    lib.vido_stream.argtypes = [vp]
    lib.vido_sync.argtypes = [vp]
    lib.vido_orb_level_info.argtypes = [vp, vp, vp, vp, vp]
    lib.vido_orb_extract.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp, C.c_int, vp]
    lib.vido_orb_extract_dev.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp, C.c_int, vp, C.c_int]
    lib.vido_bgr_to_gray_dev.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp, C.c_size_t, C.c_int]
    lib.vido_orb_get_level.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.vido_orb_get_candidates.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp]
    lib.vido_orb_describe_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, C.c_int]
    lib.vido_orb_extract_describe.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp, C.c_int, vp, vp]
    lib.vido_orb_get_blurred_level.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.vido_hamming_match.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp]
    lib.vido_hamming_match_dev.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_size_t, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int]
    lib.vido_imu_preintegrate.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp]
    lib.vido_get_kernel_times.argtypes = [vp, vp, vp, vp]
    lib.vido_track_frames.argtypes = [vp, C.POINTER(FrameInputs), C.c_int, vp, C.POINTER(TrackStats)]
    lib.vido_track_reset.argtypes = [vp]
    lib.vido_track_prefetch.argtypes = [vp, C.POINTER(FrameInputs), C.c_int]
    lib.vido_track_prefetch.restype = C.c_int
    lib.vido_map_num_frames.argtypes = [vp]
    lib.vido_map_get_poses.argtypes = [vp, vp, C.c_int]
    lib.vido_map_get_static.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int]
    lib.vido_map_get_dynamic.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int]
    lib.vido_map_get_objects.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int]
    lib.vido_map_get_dyn_tracks.argtypes = [vp, vp, vp, vp, vp, C.c_int]
    lib.vido_fba_default_params.argtypes = [C.POINTER(FbaProblem)]
    lib.vido_ba_full.argtypes = [vp, C.POINTER(FbaProblem), C.POINTER(LmStats)]
    lib.vido_full_batch.argtypes = [vp, C.POINTER(LmStats), vp]
    lib.vido_fba_save_g2o.argtypes = [C.POINTER(FbaProblem), C.c_char_p, C.c_int]
    lib.vido_convert_raw.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp]
    lib.vido_map_get_poses_rf.argtypes = [vp, vp, C.c_int]
    lib.vido_map_get_objects_rf.argtypes = [vp, C.c_int, vp, C.c_int]
    lib.vido_map_export_full_graph.argtypes = [vp] + [vp] * 14
    lib.vido_projopt_default_params.argtypes = [C.POINTER(ProjOptProblem), C.c_int]
    lib.vido_pose_opt_proj.argtypes = [vp, C.POINTER(ProjOptProblem), C.c_int, C.POINTER(LmStats)]
    lib.vido_inertial_default_params.argtypes = [C.POINTER(InertialProblem)]
    lib.vido_inertial_opt.argtypes = [vp, C.POINTER(InertialProblem), C.POINTER(LmStats)]
    lib.vido_track_set_imu.argtypes = [vp, vp, vp]
    lib.vido_metric_error.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, C.POINTER(Metric), vp]
    lib.vido_track_grab_imu.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.vido_track_get_imu_state.argtypes = [vp, C.POINTER(ImuState)]
    lib.vido_map_get_imu_frames.argtypes = [vp, vp, vp, vp, C.c_int]
    lib.vido_map_apply_scaled_rotation.argtypes = [vp, vp, C.c_float]
    lib.vido_pnp_default_params.argtypes = [C.POINTER(PnpProblem)]
    lib.vido_init_model.argtypes = [vp, C.POINTER(PnpProblem)]
    lib.vido_poseopt_default_params.argtypes = [C.POINTER(PoseOptProblem)]
    lib.vido_pose_opt_flow2.argtypes = [vp, C.POINTER(PoseOptProblem), C.c_int, C.POINTER(LmStats)]
    lib.vido_ba_default_params.argtypes = [C.POINTER(BaProblem)]
    lib.vido_ba_partial.argtypes = [vp, C.POINTER(BaProblem), C.POINTER(LmStats)]
    _lib = lib
    return lib


def default_config(**kw):
    cfg = VidoConfig()
    load_library().vido_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    return cfg


def save_g2o(path, g, n_poses, precision=0, **params):
    """the flat FullBatch graph `g` (dict keyed by FBA_KEYS) as g2o text through the C-ABI (vido_fba_save_g2o; host only) --
    the files of src/Optimizer.cc:1937,1939.  precision 0 = 6 digits like the reference's std::ostream."""
    lib = load_library()
    keep = {k: np.array(g[k], dtype=np.float32 if k in FBA_F32 else np.int32, copy=True, order="C") for k in FBA_KEYS}
    pr = FbaProblem()
    lib.vido_fba_default_params(C.byref(pr))
    pr.n_poses = n_poses
    pr.n_motions = keep["se3"].reshape(-1, 16).shape[0] - n_poses
    pr.n_points = keep["points"].reshape(-1, 3).shape[0]
    pr.n_obs, pr.n_e6, pr.n_tern = len(keep["obs_se3"]), len(keep["e6_i"]), len(keep["tern_p1"])
    for k in FBA_KEYS:
        setattr(pr, k, keep[k].ctypes.data if keep[k].size else None)
    for k, v in params.items():
        setattr(pr, k, v)
    rc = lib.vido_fba_save_g2o(C.byref(pr), str(path).encode(), precision)
    if rc != 0:
        raise VidoError(f"vido_fba_save_g2o({path}) failed: {rc}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PackedFrames:
    """a marshalled vido_frame_inputs array (Context.pack_frames)"""
    def __init__(self, arr, keep, n):
        self.arr, self.keep, self.n = arr, keep, n

    def __len__(self):
        return self.n


class Context:
    """One context per CUDA device (vido_create / vido_destroy)."""

    def __init__(self, cfg=None, **kw):
        self.lib = load_library()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self.h = self.lib.vido_create(C.byref(self.cfg))
        if not self.h:
            raise VidoError("vido_create failed: " + self.lib.vido_last_error(None).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.vido_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise VidoError(f"rc={rc}: " + self.lib.vido_last_error(self.h).decode())

    @property
    def launches(self):
        return int(self.lib.vido_kernel_launches(self.h))

    @property
    def stream(self):
        return self.lib.vido_stream(self.h)

    def sync(self):
        self._check(self.lib.vido_sync(self.h))

    def level_info(self):
        n = self.cfg.nlevels
        w = np.zeros(n, np.int32); h = np.zeros(n, np.int32); q = np.zeros(n, np.int32); s = np.zeros(n, np.float32)
        self.lib.vido_orb_level_info(self.h, _ptr(w), _ptr(h), _ptr(q), _ptr(s))
        return w, h, q, s

    def orb_extract(self, gray, cap=None):
        """gray: [H,W] or [B,H,W] uint8 numpy (host).  Returns list of keypoint arrays (KP_DTYPE)."""
        g = np.ascontiguousarray(gray, np.uint8)
        if g.ndim == 2:
            g = g[None]
        B, H, W = g.shape
        assert H == self.cfg.height and W == self.cfg.width
        cap = cap or (self.cfg.nfeatures + 64)
        out = np.zeros((B, cap), KP_DTYPE)
        n = np.zeros(B, np.int32)
        self._check(self.lib.vido_orb_extract(self.h, _ptr(g), B, H * W, W, _ptr(out), cap, _ptr(n)))
        return [out[b, :n[b]].copy() for b in range(B)]

    def orb_extract_describe(self, gray, cap=None):
        """ORBextractor::operator() with descriptors.  gray: [H,W] or [B,H,W] uint8 numpy (host).
        Returns (list of keypoint arrays, list of [n][32] uint8 descriptor arrays)."""
        g = np.ascontiguousarray(gray, np.uint8)
        if g.ndim == 2:
            g = g[None]
        B, H, W = g.shape
        assert H == self.cfg.height and W == self.cfg.width
        cap = cap or (self.cfg.nfeatures + 64)
        out = np.zeros((B, cap), KP_DTYPE)
        desc = np.zeros((B, cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        self._check(self.lib.vido_orb_extract_describe(self.h, _ptr(g), B, H * W, W, _ptr(out), cap, _ptr(n), _ptr(desc)))
        return [out[b, :n[b]].copy() for b in range(B)], [desc[b, :n[b]].copy() for b in range(B)]

    def orb_describe_dev(self, d_kps_ptr, d_nkp_ptr, nframes, cap, d_desc_ptr, sync=False):
        self._check(self.lib.vido_orb_describe_dev(self.h, d_kps_ptr, d_nkp_ptr, nframes, cap, d_desc_ptr, 1 if sync else 0))

    def get_blurred_level(self, frame, level):
        w, h, _, _ = self.level_info()
        out = np.zeros((h[level], w[level]), np.uint8)
        self._check(self.lib.vido_orb_get_blurred_level(self.h, frame, level, _ptr(out)))
        return out

    def hamming_match(self, query, train):
        """brute-force Hamming matching of [n][32] uint8 descriptor arrays (host): best train index, its distance, second distance"""
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        bi = np.zeros(len(q), np.int32); bd = np.zeros(len(q), np.int32); sd = np.zeros(len(q), np.int32)
        self._check(self.lib.vido_hamming_match(self.h, _ptr(q), len(q), _ptr(t), len(t), _ptr(bi), _ptr(bd), _ptr(sd)))
        return bi, bd, sd

    def hamming_match_dev(self, d_q, q_stride, d_nq, d_t, t_stride, d_nt, npairs, qcap, d_best_idx, d_best_dist, d_second, sync=False):
        self._check(self.lib.vido_hamming_match_dev(self.h, d_q, q_stride, d_nq, d_t, t_stride, d_nt, npairs, qcap, d_best_idx,
                                                    d_best_dist, d_second, 1 if sync else 0))

    def orb_extract_dev(self, d_gray_ptr, nframes, frame_stride, stride, d_out_ptr, cap, d_n_ptr, sync=False):
        self._check(self.lib.vido_orb_extract_dev(self.h, d_gray_ptr, nframes, frame_stride, stride, d_out_ptr, cap,
                                                  d_n_ptr, 1 if sync else 0))

    def bgr_to_gray_dev(self, d_bgr, nframes, frame_stride, stride, d_gray, gray_frame_stride, gray_stride):
        self._check(self.lib.vido_bgr_to_gray_dev(self.h, d_bgr, nframes, frame_stride, stride, d_gray,
                                                  gray_frame_stride, gray_stride))

    def get_level(self, frame, level):
        w, h, _, _ = self.level_info()
        out = np.zeros((h[level], w[level]), np.uint8)
        self._check(self.lib.vido_orb_get_level(self.h, frame, level, _ptr(out)))
        return out

    def get_candidates(self, frame, level, cap=1 << 18):
        xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
        n = np.zeros(1, np.int32)
        self._check(self.lib.vido_orb_get_candidates(self.h, frame, level, _ptr(xs), _ptr(ys), _ptr(sc), cap, _ptr(n)))
        return xs[:n[0]].copy(), ys[:n[0]].copy(), sc[:n[0]].copy()

    def ba_partial(self, poses, rel_motion, points, obs_pose, obs_point, obs_xyz, **params):
        """Sliding-window graph optimisation (vido_ba_partial).  Inputs are copied; returns
        (poses, rel_motion, points, iterations, stats)."""
        poses = np.ascontiguousarray(poses, np.float32).copy()
        rel = np.ascontiguousarray(rel_motion, np.float32).copy()
        pts = np.ascontiguousarray(points, np.float32).copy()
        op = np.ascontiguousarray(obs_pose, np.int32)
        ol = np.ascontiguousarray(obs_point, np.int32)
        ox = np.ascontiguousarray(obs_xyz, np.float32)
        pr = BaProblem()
        self.lib.vido_ba_default_params(C.byref(pr))
        pr.n_poses, pr.n_points, pr.n_obs = len(poses), len(pts), len(op)
        pr.poses, pr.rel_motion, pr.points = _ptr(poses), _ptr(rel), _ptr(pts)
        pr.obs_pose, pr.obs_point, pr.obs_xyz = _ptr(op), _ptr(ol), _ptr(ox)
        for k, v in params.items():
            setattr(pr, k, v)
        st = LmStats()
        self._check(self.lib.vido_ba_partial(self.h, C.byref(pr), C.byref(st)))
        return poses, rel, pts, st.iterations, st

    def pose_opt_flow2(self, problems, want_stats=True):
        """problems: list of dicts(obs, flow, depth, Tcw_init, Tcw_last, K, [params...]).  One launch for all.
        Returns a list of (Tcw 4x4 f32, refined flow [n,2], inlier [n], n_inliers, [LmStats per round])."""
        m = len(problems)
        arr = (PoseOptProblem * m)()
        keep = []
        for k, d in enumerate(problems):
            pr = arr[k]
            self.lib.vido_poseopt_default_params(C.byref(pr))
            obs = np.ascontiguousarray(d["obs"], np.float32); fl = np.ascontiguousarray(d["flow"], np.float32)
            dep = np.ascontiguousarray(d["depth"], np.float32)
            n = len(obs)
            fo = np.zeros((n, 2), np.float32); inl = np.zeros(n, np.int32)
            keep.append((obs, fl, dep, fo, inl))
            pr.n = n
            pr.obs_xy, pr.flow_xy, pr.depth, pr.flow_out, pr.inlier = _ptr(obs), _ptr(fl), _ptr(dep), _ptr(fo), _ptr(inl)
            pr.Tcw_init[:] = np.asarray(d["Tcw_init"], np.float32).reshape(-1).tolist()
            pr.Tcw_last[:] = np.asarray(d["Tcw_last"], np.float32).reshape(-1).tolist()
            pr.fx, pr.fy, pr.cx, pr.cy = [float(v) for v in d["K"]]
            for key, v in d.get("params", {}).items():
                setattr(pr, key, v)
        stats = (LmStats * (4 * m))() if want_stats else None
        self._check(self.lib.vido_pose_opt_flow2(self.h, arr, m, stats))
        out = []
        for k in range(m):
            pr = arr[k]
            st = [stats[4 * k + r] for r in range(pr.rounds)] if want_stats else None
            out.append((np.array(pr.Tcw_out[:], np.float32).reshape(4, 4), keep[k][3], keep[k][4], int(pr.n_inliers), st))
        return out

    def init_model(self, cur_xy, pts3d, valid, Tcw_motion, K, **params):
        """vido_init_model: returns (Tcw 4x4 f32, inlier ids, winner, ransac_inliers, mm_inliers)"""
        pr = PnpProblem()
        self.lib.vido_pnp_default_params(C.byref(pr))
        cur = np.ascontiguousarray(cur_xy, np.float32); pts = np.ascontiguousarray(pts3d, np.float32)
        val = np.ascontiguousarray(valid, np.int32) if valid is not None else None
        ids = np.zeros(max(len(cur), 1), np.int32)
        pr.n = len(cur)
        pr.cur_xy, pr.pts3d, pr.inlier_ids = _ptr(cur), _ptr(pts), _ptr(ids)
        pr.valid = _ptr(val) if val is not None else None
        pr.Tcw_motion[:] = np.asarray(Tcw_motion, np.float32).reshape(-1).tolist()
        pr.fx, pr.fy, pr.cx, pr.cy = [float(v) for v in K]
        for k, v in params.items():
            setattr(pr, k, v)
        self._check(self.lib.vido_init_model(self.h, C.byref(pr)))
        return (np.array(pr.Tcw_out[:], np.float32).reshape(4, 4), ids[:pr.n_inliers].copy(), pr.winner,
                pr.ransac_inliers, pr.mm_inliers)

    # ---- per-frame driver
    def pack_frames(self, frames):
        """marshal a list of frame dicts once (vido_frame_inputs array); the result can be passed to track_frames /
        track_prefetch any number of times -- a streaming caller then pays no Python marshalling per call"""
        arr, keep = self._frame_inputs(frames)
        return PackedFrames(arr, keep, len(frames))

    def _frame_inputs(self, frames):
        if isinstance(frames, PackedFrames):
            return frames.arr, frames.keep
        n = len(frames)
        arr = (FrameInputs * n)()
        keep = []
        for k, f in enumerate(frames):
            fi = arr[k]
            if f.get("on_device"):
                fi.image, fi.depth, fi.flow, fi.mask = f["image"], f["depth"], f["flow"], f["mask"]
                fi.channels, fi.on_device = int(f["channels"]), 1
            elif f.get("host_ptrs"):  # raw host pointers (e.g. pinned torch tensors): image, depth, flow, mask + channels
                fi.image, fi.depth, fi.flow, fi.mask = f["image"], f["depth"], f["flow"], f["mask"]
                fi.channels, fi.on_device = int(f["channels"]), 0
            else:
                img = np.ascontiguousarray(f["image"], np.uint8); dep = np.ascontiguousarray(f["depth"], np.float32)
                flo = np.ascontiguousarray(f["flow"], np.float32); msk = np.ascontiguousarray(f["mask"], np.int32)
                keep.append((img, dep, flo, msk))
                fi.image, fi.depth, fi.flow, fi.mask = _ptr(img), _ptr(dep), _ptr(flo), _ptr(msk)
                fi.channels, fi.on_device = (1 if img.ndim == 2 else img.shape[2]), 0
            fi.write_back_depth = int(f.get("write_back_depth", 0))
            fi.timestamp = float(f.get("timestamp", 0.1 * k))
        return arr, keep

    # ---- VIO mode (sensor = IMU_RGBD)
    def convert_raw(self, nframes, d_bgr=None, d_depth=None, d_mask=None, bayer=None, depth16=None, mask8=None):
        """vido_convert_raw: raw Bayer RG / u16 depth / u8 mask (numpy host arrays or integer device pointers) -> device
        buffers (integer device pointers) in the tracker's input types; asynchronous on the context's stream"""
        def src(a):
            if a is None:
                return None, None
            if isinstance(a, int):
                return a, None
            a = np.ascontiguousarray(a)
            return a.ctypes.data, a
        pb, kb = src(bayer); pd, kd = src(depth16); pm, km = src(mask8)
        self._check(self.lib.vido_convert_raw(self.h, pb, pd, pm, nframes, d_bgr, d_depth, d_mask))
        self.sync()   # the host sources (kb, kd, km) may go away after the call
        return kb, kd, km

    def track_set_imu(self, Tbc, noise):
        """Tracking::ParseIMUParamFile: Tbc 4x4, noise (ng, na, ngw, naw) as given to IMU::Calib"""
        if Tbc is None:   # back to sensor = RGBD
            self._check(self.lib.vido_track_set_imu(self.h, None, None))
            return
        T = np.ascontiguousarray(Tbc, np.float32).reshape(16); nz = np.ascontiguousarray(noise, np.float32)
        self._check(self.lib.vido_track_set_imu(self.h, _ptr(T), _ptr(nz)))

    def track_grab_imu(self, samples, frames_ahead=0):
        """Tracking::GrabImuData for the frame `frames_ahead` frames after the next one to be tracked"""
        s = np.ascontiguousarray(samples, IMU_SAMPLE)
        self._check(self.lib.vido_track_grab_imu(self.h, _ptr(s), len(s), int(frames_ahead)))

    def imu_state(self):
        st = ImuState()
        self._check(self.lib.vido_track_get_imu_state(self.h, C.byref(st)))
        return st

    def map_imu_frames(self):
        n = self.lib.vido_map_get_imu_frames(self.h, None, None, None, 0)
        T = np.zeros((max(n, 0), 16), np.float32); v = np.zeros((max(n, 0), 3), np.float32); b = np.zeros((max(n, 0), 6), np.float32)
        if n > 0:
            self.lib.vido_map_get_imu_frames(self.h, _ptr(T), _ptr(v), _ptr(b), n)
        return T.reshape(-1, 4, 4), v, b

    def metric_error(self, cam_gt, obj_pose_pre=None, obj_motion_gt=None, refined=False):
        """Tracking::GetMetricError on the context's Map; returns (Metric, per-item (t, r) array)"""
        g = np.ascontiguousarray(cam_gt, np.float32).reshape(-1, 16)
        no = 0 if obj_pose_pre is None else len(obj_pose_pre)
        pp = np.ascontiguousarray(obj_pose_pre, np.float32).reshape(-1, 16) if no else None
        mg = np.ascontiguousarray(obj_motion_gt, np.float32).reshape(-1, 16) if no else None
        m = Metric()
        per = np.zeros((max(len(g) - 1, 0) + no, 2), np.float32)
        self._check(self.lib.vido_metric_error(self.h, _ptr(g), len(g), int(refined), _ptr(pp) if no else None,
                                               _ptr(mg) if no else None, no, C.byref(m), _ptr(per)))
        return m, per

    def map_apply_scaled_rotation(self, R, s):
        """Map::ApplyScaledRotation(R, s)"""
        Rm = np.ascontiguousarray(R, np.float32).reshape(9)
        self._check(self.lib.vido_map_apply_scaled_rotation(self.h, _ptr(Rm), float(s)))

    def track_frames(self, frames, want_stats=True, imu=None):
        """frames: list of dicts {image, depth, flow, mask} holding either numpy host arrays or integer device
        pointers (then pass on_device=True and channels).  imu (VIO mode): per frame the IMU samples System::TrackRGBD would
        receive with it.  Returns (Tcw [n,4,4] f32, [stats dict] or None)."""
        n = len(frames)
        if imu is not None:
            for k, smp in enumerate(imu):
                self.track_grab_imu(smp, k)
        arr, keep = self._frame_inputs(frames)
        T = np.zeros((n, 16), np.float32)
        st = (TrackStats * n)() if want_stats else None
        self._check(self.lib.vido_track_frames(self.h, arr, n, _ptr(T), st))
        return T.reshape(n, 4, 4), ([s.as_dict() for s in st] if want_stats else None)

    def track_prefetch(self, frames):
        """Hint: start the host->device copy of the frames the NEXT track_frames call will pass (same buffers)."""
        arr, keep = self._frame_inputs(frames)
        self._prefetch_keep = keep  # the host arrays must outlive the copy
        self._check(self.lib.vido_track_prefetch(self.h, arr, len(frames)))

    def track_reset(self):
        self._check(self.lib.vido_track_reset(self.h))

    def map_poses(self):
        n = self.lib.vido_map_num_frames(self.h)
        P = np.zeros((max(n, 1), 16), np.float32)
        self.lib.vido_map_get_poses(self.h, _ptr(P), n)
        return P[:n].reshape(n, 4, 4)

    def map_static(self, frame, cap=4096):
        xy = np.zeros((cap, 2), np.float32); dep = np.zeros(cap, np.float32); p3 = np.zeros((cap, 3), np.float32)
        asso = np.zeros(cap, np.int32)
        n = self.lib.vido_map_get_static(self.h, frame, _ptr(xy), _ptr(dep), _ptr(p3), _ptr(asso), cap)
        if n < 0:
            raise VidoError("bad frame index")
        return xy[:n].copy(), dep[:n].copy(), p3[:n].copy(), asso[:n].copy()

    def map_dynamic(self, frame, cap=32768):
        """Map::vpFeatDyn / vfDepDyn / vp3DPointDyn / vnAssoDyn / vnFeatLabel of one frame"""
        xy = np.zeros((cap, 2), np.float32); dep = np.zeros(cap, np.float32); p3 = np.zeros((cap, 3), np.float32)
        asso = np.zeros(cap, np.int32); lab = np.zeros(cap, np.int32)
        n = self.lib.vido_map_get_dynamic(self.h, frame, _ptr(xy), _ptr(dep), _ptr(p3), _ptr(asso), _ptr(lab), cap)
        if n < 0:
            raise VidoError("bad frame index")
        return xy[:n].copy(), dep[:n].copy(), p3[:n].copy(), asso[:n].copy(), lab[:n].copy()

    def map_objects(self, frame, cap=64):
        """(tracking label, semantic label, motion 4x4, centre) of the objects with an estimated motion in `frame` (>= 1)"""
        lab = np.zeros(cap, np.int32); sem = np.zeros(cap, np.int32); mot = np.zeros((cap, 16), np.float32)
        cen = np.zeros((cap, 3), np.float32)
        n = max(self.lib.vido_map_get_objects(self.h, frame, _ptr(lab), _ptr(sem), _ptr(mot), _ptr(cen), cap), 0)
        return lab[:n].copy(), sem[:n].copy(), mot[:n].reshape(n, 4, 4).copy(), cen[:n].copy()

    def map_dyn_tracks(self, cap=1 << 20):
        ln = np.zeros(cap, np.int32); oid = np.zeros(cap, np.int32); ff = np.zeros(cap, np.int32); fj = np.zeros(cap, np.int32)
        n = self.lib.vido_map_get_dyn_tracks(self.h, _ptr(ln), _ptr(oid), _ptr(ff), _ptr(fj), cap)
        return ln[:n].copy(), oid[:n].copy(), ff[:n].copy(), fj[:n].copy()

    # ---- full-sequence optimisation (Optimizer::FullBatchOptimization)
    def ba_full(self, g, n_poses, **params):
        """solve a flat graph (dict keyed by FBA_KEYS); returns (se3 [n,4,4], points [m,3], LmStats)"""
        keep = {k: np.array(g[k], dtype=np.float32 if k in FBA_F32 else np.int32, copy=True, order="C") for k in FBA_KEYS}
        pr = FbaProblem()
        self.lib.vido_fba_default_params(C.byref(pr))
        pr.n_poses = n_poses
        pr.n_motions = keep["se3"].reshape(-1, 16).shape[0] - n_poses
        pr.n_points = keep["points"].reshape(-1, 3).shape[0]
        pr.n_obs, pr.n_e6, pr.n_tern = len(keep["obs_se3"]), len(keep["e6_i"]), len(keep["tern_p1"])
        for k in FBA_KEYS:
            setattr(pr, k, _ptr(keep[k]).value if keep[k].size else None)
        for k, v in params.items():
            setattr(pr, k, v)
        st = LmStats()
        self._check(self.lib.vido_ba_full(self.h, C.byref(pr), C.byref(st)))
        return keep["se3"].reshape(-1, 4, 4), keep["points"].reshape(-1, 3), st

    def full_batch(self):
        """FullBatchOptimization on the context's map; returns (LmStats, sizes[6])"""
        st = LmStats(); sizes = np.zeros(6, np.int32)
        self._check(self.lib.vido_full_batch(self.h, C.byref(st), _ptr(sizes)))
        return st, sizes

    def map_poses_rf(self):
        n = self.lib.vido_map_num_frames(self.h)
        P = np.zeros((max(n, 1), 16), np.float32)
        self.lib.vido_map_get_poses_rf(self.h, _ptr(P), n)
        return P[:n].reshape(n, 4, 4)

    def map_objects_rf(self, frame, cap=64):
        mot = np.zeros((cap, 16), np.float32)
        n = max(self.lib.vido_map_get_objects_rf(self.h, frame, _ptr(mot), cap), 0)
        return mot[:n].reshape(n, 4, 4).copy()

    def export_full_graph(self):
        """flat FullBatch graph ("keyframe factors") of the map: (dict keyed by FBA_KEYS, n_poses)"""
        sizes = np.zeros(6, np.int32)
        self._check(self.lib.vido_map_export_full_graph(self.h, _ptr(sizes), *([None] * 13)))
        npo, nmo, npt, nob, ne6, nte = [int(v) for v in sizes]
        g = dict(se3=np.zeros((npo + nmo, 16), np.float32), points=np.zeros((npt, 3), np.float32),
                 e6_i=np.zeros(ne6, np.int32), e6_j=np.zeros(ne6, np.int32), e6_kind=np.zeros(ne6, np.int32),
                 e6_meas=np.zeros((ne6, 16), np.float32), obs_se3=np.zeros(nob, np.int32), obs_point=np.zeros(nob, np.int32),
                 obs_kind=np.zeros(nob, np.int32), obs_xyz=np.zeros((nob, 3), np.float32), tern_p1=np.zeros(nte, np.int32),
                 tern_p2=np.zeros(nte, np.int32), tern_h=np.zeros(nte, np.int32))
        self._check(self.lib.vido_map_export_full_graph(self.h, _ptr(sizes), *[_ptr(g[k]) if g[k].size else None for k in FBA_KEYS]))
        return g, npo

    def pose_opt_proj(self, problems):
        """problems: list of dicts(kind, obs_xy, pts3d, T_init, K=(fx,fy,cx,cy) | P=3x4, + optional rp_thres / its).
        One launch for all.  Returns [(T 4x4, inlier flags, LmStats)]"""
        n = len(problems)
        arr = (ProjOptProblem * n)()
        st = (LmStats * n)()
        keep = []
        for k, d in enumerate(problems):
            pr = arr[k]
            self.lib.vido_projopt_default_params(C.byref(pr), int(d["kind"]))
            obs = np.ascontiguousarray(d["obs_xy"], np.float32).reshape(-1, 2)
            pts = np.ascontiguousarray(d["pts3d"], np.float32).reshape(-1, 3)
            inl = np.zeros(len(obs), np.int32)
            keep.append((obs, pts, inl))
            pr.n = len(obs)
            pr.obs_xy, pr.pts3d, pr.inlier = _ptr(obs).value, _ptr(pts).value, _ptr(inl).value
            pr.T_init[:] = [float(v) for v in np.asarray(d["T_init"], np.float32).reshape(16)]
            if d.get("K") is not None:
                pr.fx, pr.fy, pr.cx, pr.cy = d["K"]
            if d.get("P") is not None:
                pr.P[:] = [float(v) for v in np.asarray(d["P"], np.float64).reshape(12)]
            for key in ("rp_thres", "its"):
                if key in d:
                    setattr(pr, key, d[key])
        self._check(self.lib.vido_pose_opt_proj(self.h, arr, n, st))
        return [(np.array(arr[k].T_out[:], np.float32).reshape(4, 4), keep[k][2].copy(), st[k]) for k in range(n)]

    def inertial_opt(self, Rwb, twb, vel, preint, bias_lin, Rwg, scale=1.0, bg=(0, 0, 0), ba=(0, 0, 0), **params):
        """Optimizer::InertialOptimization; returns dict(velocity, Rwg, scale, bg, ba, stats)"""
        keep = [np.ascontiguousarray(Rwb, np.float32).reshape(-1, 9), np.ascontiguousarray(twb, np.float32).reshape(-1, 3),
                np.array(vel, np.float32).reshape(-1, 3).copy(), np.ascontiguousarray(preint, IMU_PREINT),
                np.ascontiguousarray(bias_lin, np.float32).reshape(-1, 6)]
        pr = InertialProblem()
        self.lib.vido_inertial_default_params(C.byref(pr))
        pr.n_frames = keep[0].shape[0]
        pr.Rwb, pr.twb, pr.velocity, pr.preint, pr.bias_lin = [_ptr(a).value for a in keep]
        pr.Rwg[:] = [float(v) for v in np.asarray(Rwg, np.float64).reshape(9)]
        pr.scale = float(scale)
        pr.bg[:] = [float(v) for v in bg]
        pr.ba[:] = [float(v) for v in ba]
        for k, v in params.items():
            setattr(pr, k, v)
        st = LmStats()
        self._check(self.lib.vido_inertial_opt(self.h, C.byref(pr), C.byref(st)))
        return dict(velocity=keep[2], Rwg=np.array(pr.Rwg[:]).reshape(3, 3), scale=pr.scale, bg=np.array(pr.bg[:]),
                    ba=np.array(pr.ba[:]), stats=st)

    def kernel_times(self):
        """device ms / timed regions of (ORB front-end, init model, pose optimisation, window BA) + BA algorithmic bytes"""
        ms = np.zeros(4, np.float64); n = np.zeros(4, np.int64); b = np.zeros(1, np.float64)
        self._check(self.lib.vido_get_kernel_times(self.h, _ptr(ms), _ptr(n), _ptr(b)))
        return ms, n, float(b[0])

    def imu_preintegrate(self, samples, t_prev, t_cur, bias, noise):
        """batched IMU preintegration: t_prev/t_cur arrays of length njobs, bias [njobs,6]; returns IMU_PREINT array"""
        s = np.ascontiguousarray(samples, IMU_SAMPLE)
        tp = np.ascontiguousarray(np.atleast_1d(t_prev), np.float64); tc = np.ascontiguousarray(np.atleast_1d(t_cur), np.float64)
        nj = len(tp)
        b = np.ascontiguousarray(np.broadcast_to(np.asarray(bias, np.float32).reshape(-1, 6), (nj, 6)))
        nz = np.ascontiguousarray(noise, np.float32)
        out = np.zeros(nj, IMU_PREINT)
        self._check(self.lib.vido_imu_preintegrate(self.h, _ptr(s), len(s), _ptr(tp), _ptr(tc), nj, _ptr(b), _ptr(nz), _ptr(out)))
        return out


# ---------------------------------------------------------------------------------------------------------------
# on-disk input formats of the demo (host/InputDecode.h, built into libvido_slam.so)
_io = None


def _io_lib():
    global _io
    if _io is None:
        load_library()   # libvido_slam.so links libvido_b200.so
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvido_slam.so")
        if not os.path.exists(path):
            raise VidoError(f"{path} is missing: run __graft_entry__.build()")
        _io = C.CDLL(path)
        _io.vido_io_read_png.argtypes = [C.c_char_p] + [C.POINTER(C.c_int32)] * 4 + [C.c_void_p, C.c_size_t]
        _io.vido_io_read_flo.argtypes = [C.c_char_p] + [C.POINTER(C.c_int32)] * 2 + [C.c_void_p, C.c_size_t]
        _io.vido_io_load_kaist_imu.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        _io.vido_io_load_kaist_timestamps.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
    return _io


def read_png(path):
    """cv::imread(path, IMREAD_UNCHANGED) for 8 / 16-bit grey, grey-alpha, RGB(A) PNGs (channels in FILE order, i.e. RGB)"""
    lib = _io_lib()
    w, h, ch, bd = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    if lib.vido_io_read_png(str(path).encode(), C.byref(w), C.byref(h), C.byref(ch), C.byref(bd), None, 0) != 0:
        raise VidoError(f"cannot decode {path}")
    a = np.zeros((h.value, w.value, ch.value), np.uint8 if bd.value == 8 else np.uint16)
    if lib.vido_io_read_png(str(path).encode(), None, None, None, None, a.ctypes.data, a.nbytes) != 0:
        raise VidoError(f"cannot decode {path}")
    return a[:, :, 0] if ch.value == 1 else a


def read_flo(path):
    """cv::optflow::readOpticalFlow: [H, W, 2] float32"""
    lib = _io_lib()
    w, h = C.c_int32(), C.c_int32()
    if lib.vido_io_read_flo(str(path).encode(), C.byref(w), C.byref(h), None, 0) != 0:
        raise VidoError(f"cannot read {path}")
    a = np.zeros((h.value, w.value, 2), np.float32)
    if lib.vido_io_read_flo(str(path).encode(), None, None, a.ctypes.data, a.size) != 0:
        raise VidoError(f"cannot read {path}")
    return a


def load_kaist_imu(path, cap=1 << 21):
    """LoadIMU of the demo: rows (t [s], ax, ay, az, wx, wy, wz)"""
    a = np.zeros((cap, 7), np.float64)
    n = _io_lib().vido_io_load_kaist_imu(str(path).encode(), a.ctypes.data, cap)
    if n < 0:
        raise VidoError(f"cannot read {path}")
    return a[:n].copy()


def load_kaist_timestamps(image_dir, cap=1 << 18):
    """LoadKaistImg of the demo: (file stems [19 characters], times [s])"""
    t = np.zeros(cap, np.float64)
    names = C.create_string_buffer(20 * cap)
    n = _io_lib().vido_io_load_kaist_timestamps(str(image_dir).encode(), t.ctypes.data, names, cap)
    if n < 0:
        raise VidoError(f"cannot read {image_dir}/../vTimestampsImage.txt")
    return [names.raw[20 * i:20 * i + 20].split(b"\0")[0].decode() for i in range(n)], t[:n].copy()
