"""g2o text files of the FullBatch graph -- what `optimizer.save("dynamic_slam_graph_before_opt.g2o")` /
`..."after_opt.g2o"` write in Optimizer::FullBatchOptimization (src/Optimizer.cc:1937,1939).

Layout (g2o/core/optimizable_graph.cpp:589-622, :817-860; parameter_container.cpp:86-95): the parameters first, then every
vertex `TAG id estimate` (a `FIX id` line after a fixed one), then every edge `TAG vertex-ids payload`.  Tags and payloads of
the five element types of this graph (g2o/types/types_slam3d.cpp:37-45):

  PARAMS_SE3OFFSET id  x y z qx qy qz qw                                  parameter_se3_offset.cpp:59-64 (the camera offset, id 0)
  VERTEX_SE3:QUAT id   x y z qx qy qz qw                                  vertex_se3.cpp:58-64, isometry3d_mappings.cpp:109-116
  VERTEX_TRACKXYZ id   x y z                                              vertex_pointxyz.cpp:47-53
  EDGE_SE3_PRIOR v  0  x y z qx qy qz qw  info[21 upper triangle]         edge_se3_prior.cpp:77-86   (src/Optimizer.cc:1369-1376)
  EDGE_SE3:QUAT i j    x y z qx qy qz qw  info[21]                        edge_se3.cpp:67-75        (:1389-1400, :1611-1638)
  EDGE_SE3_TRACKXYZ v p  0  x y z  info[6]                                edge_se3_pointxyz.cpp:88-96 (:1425-1436, :1485-1520)
  EDGE_SE3_MOTION p1 p2 h  0 0 0  info[6]                                 types_dyn_slam3d.cpp:44-51  (:1733-1745)

Robust kernels are not part of the format (g2o does not save them).  Numbers are printed like a default std::ostream
(6 significant digits) unless `precision` says otherwise; g2o reads any precision.  Vertex ids start at 1 like the
reference's counter (:1353): SE3 vertices (camera poses, then object motions) first, then the points -- the reference
numbers them in creation order frame by frame, which the flat graph (vido_fba_problem) does not record; no result of g2o
depends on the ids.

The writer is also how the graph is checked against the reference's own g2o build: tests/golden/make_g2o_golden.py loads
these files with /root/reference/vido_slam/3rdparty/g2o/lib/libg2o.so and records the chi2 its edge classes compute.
"""
import numpy as np

FIRST_ID = 1
SIGMA_DEFAULTS = dict(sigma2_cam=0.0001, sigma2_3d_sta=80.0, sigma2_3d_dyn=80.0, sigma2_obj=100.0, sigma2_smooth=0.001,
                      prior_info=100000.0)  # src/Optimizer.cc:1290-1295, :1373


def _quat(R):
    """unit quaternion (x, y, z, w) of a rotation matrix, w >= 0 (Shepperd's selection of the largest component)"""
    R = np.asarray(R, np.float64)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    q /= np.linalg.norm(q)
    return q if q[3] >= 0 else -q


def _rot(q):
    x, y, z, w = np.asarray(q, np.float64) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _qt(T):
    T = np.asarray(T, np.float64).reshape(4, 4)
    return np.concatenate([T[:3, 3], _quat(T[:3, :3])])


def _upper(scale, n):
    """upper triangle, row by row, of scale * I_n"""
    return [scale if i == j else 0.0 for i in range(n) for j in range(i, n)]


def write_g2o(path, g, n_poses, precision=6, **sigmas):
    """Write the flat FullBatch graph `g` (dict keyed by FBA_KEYS, see Context.export_full_graph) as g2o text.
    Information matrices are the reference's: I / sigma2 with the float32 sigma2 of src/Optimizer.cc:1290-1295."""
    s = dict(SIGMA_DEFAULTS, **sigmas)
    inv = {k: 1.0 / float(np.float32(v)) for k, v in s.items() if k != "prior_info"}
    se3 = np.asarray(g["se3"], np.float64).reshape(-1, 4, 4)
    pts = np.asarray(g["points"], np.float64).reshape(-1, 3)
    n_se3 = se3.shape[0]
    sid = lambda k: FIRST_ID + int(k)
    pid = lambda k: FIRST_ID + n_se3 + int(k)
    fmt = lambda vals: " ".join(format(float(v), ".%dg" % precision) for v in vals)
    out = ["PARAMS_SE3OFFSET 0 " + fmt([0, 0, 0, 0, 0, 0, 1]) + " "]
    for k in range(n_se3):
        out.append("VERTEX_SE3:QUAT %d %s " % (sid(k), fmt(_qt(se3[k]))))
    for k in range(pts.shape[0]):
        out.append("VERTEX_TRACKXYZ %d %s " % (pid(k), fmt(pts[k])))
    if n_poses > 0:
        out.append("EDGE_SE3_PRIOR %d 0 %s %s " % (sid(0), fmt(_qt(se3[0])), fmt(_upper(float(np.float32(s["prior_info"])), 6))))
    meas = np.asarray(g["e6_meas"], np.float64).reshape(-1, 4, 4)
    for e in range(len(g["e6_i"])):
        w = inv["sigma2_cam"] if int(g["e6_kind"][e]) == 0 else inv["sigma2_smooth"]
        out.append("EDGE_SE3:QUAT %d %d %s %s " % (sid(g["e6_i"][e]), sid(g["e6_j"][e]), fmt(_qt(meas[e])), fmt(_upper(w, 6))))
    xyz = np.asarray(g["obs_xyz"], np.float64).reshape(-1, 3)
    for e in range(len(g["obs_se3"])):
        w = inv["sigma2_3d_sta"] if int(g["obs_kind"][e]) == 0 else inv["sigma2_3d_dyn"]
        out.append("EDGE_SE3_TRACKXYZ %d %d 0 %s %s " % (sid(g["obs_se3"][e]), pid(g["obs_point"][e]), fmt(xyz[e]), fmt(_upper(w, 3))))
    for e in range(len(g["tern_p1"])):
        out.append("EDGE_SE3_MOTION %d %d %d %s %s " % (pid(g["tern_p1"][e]), pid(g["tern_p2"][e]), sid(g["tern_h"][e]),
                                                       fmt([0, 0, 0]), fmt(_upper(inv["sigma2_obj"], 3))))
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    return len(out)


def read_g2o(path):
    """Parse a g2o text file with the seven line types above (as written by write_g2o or by g2o's own save()).
    Returns dict(se3 {id: 4x4}, points {id: xyz}, fixed [ids], prior [(v, T, info)], e6 [(i, j, T, info6x6)],
    obs [(v, p, xyz, info3x3)], tern [(p1, p2, h, meas, info3x3)], offset {id: 4x4})."""
    def T_of(v):
        T = np.eye(4)
        T[:3, :3] = _rot(v[3:7])
        T[:3, 3] = v[:3]
        return T

    def sym(v, n):
        M = np.zeros((n, n))
        it = iter(v)
        for i in range(n):
            for j in range(i, n):
                M[i, j] = M[j, i] = next(it)
        return M

    G = dict(se3={}, points={}, fixed=[], prior=[], e6=[], obs=[], tern=[], offset={})
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            tag, v = tok[0], tok[1:]
            f = lambda a: np.array([float(x) for x in a])
            if tag == "PARAMS_SE3OFFSET":
                G["offset"][int(v[0])] = T_of(f(v[1:8]))
            elif tag == "VERTEX_SE3:QUAT":
                G["se3"][int(v[0])] = T_of(f(v[1:8]))
            elif tag == "VERTEX_TRACKXYZ":
                G["points"][int(v[0])] = f(v[1:4])
            elif tag == "FIX":
                G["fixed"].append(int(v[0]))
            elif tag == "EDGE_SE3_PRIOR":
                G["prior"].append((int(v[0]), T_of(f(v[2:9])), sym(f(v[9:30]), 6)))
            elif tag == "EDGE_SE3:QUAT":
                G["e6"].append((int(v[0]), int(v[1]), T_of(f(v[2:9])), sym(f(v[9:30]), 6)))
            elif tag == "EDGE_SE3_TRACKXYZ":
                G["obs"].append((int(v[0]), int(v[1]), f(v[3:6]), sym(f(v[6:12]), 3)))
            elif tag == "EDGE_SE3_MOTION":
                G["tern"].append((int(v[0]), int(v[1]), int(v[2]), f(v[3:6]), sym(f(v[6:12]), 3)))
            else:
                raise ValueError("unknown g2o tag " + tag)
    return G
