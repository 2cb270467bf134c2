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
        # PoseOptimizationFlow2Cam (:1133): the last frame's features, their flow and depth; refined flow replaces the match
        T1, fo, inl, ninl, _ = ol.poseopt_flow2cam(last["stat_keys"][TM], last["flow_next"][TM], last["stat_depth"][TM], T0, last["Tcw"], self.K)
        if len(TM) >= 3:
            for i, k in enumerate(TM):
                if inl[i]:
                    sk[k, 0] = F(np.float64(last["stat_keys"][k, 0]) + np.float64(fo[i, 0]))
                    sk[k, 1] = F(np.float64(last["stat_keys"][k, 1]) + np.float64(fo[i, 1]))
                else:
                    TM[i] = -1
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
