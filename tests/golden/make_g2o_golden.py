"""Generates tests/golden/g2o_golden.npz: a small FullBatch graph (from the oracle tracker on a seeded synthetic sequence with
moving objects), before and after the oracle's own FullBatch optimisation, together with the chi2 that THE REFERENCE'S OWN g2o
BUILD (/root/reference/vido_slam/3rdparty/g2o/lib/libg2o.so, loaded by oracle/_ref/g2o_chi2 -- `make -C oracle g2o_ref`)
computes for the g2o text files vido-slam_b200/g2o_text.py writes: total and per edge type (prior, odometry, smoothness,
static / dynamic point observations, landmark motion).  Also stores a tiny graph exactly as g2o's own save() printed it after
loading our file (tests/golden/g2o_saved_by_reference.g2o) for the reader / writer format test.

Runs in the authoring container only (needs /root/reference); the fixture travels.  usage: python tests/golden/make_g2o_golden.py
"""
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as ol  # noqa: E402
import synth  # noqa: E402

g2o_text = importlib.import_module("vido-slam_b200.g2o_text")
EDGE_KEYS = ("e6_i", "e6_j", "e6_kind", "e6_meas", "obs_se3", "obs_point", "obs_kind", "obs_xyz", "tern_p1", "tern_p2", "tern_h")
SUBSETS = ("prior", "odometry", "smoothness", "obs_static", "obs_dynamic", "motion")


def subset(g, which):
    """the graph with the edges of one type only (the prior is always written: 'prior' keeps nothing else)"""
    s = {k: np.array(g[k], copy=True) for k in g}
    e6 = np.zeros(len(g["e6_i"]), bool); ob = np.zeros(len(g["obs_se3"]), bool); te = np.zeros(len(g["tern_p1"]), bool)
    if which == "odometry": e6 = np.asarray(g["e6_kind"]) == 0
    if which == "smoothness": e6 = np.asarray(g["e6_kind"]) == 1
    if which == "obs_static": ob = np.asarray(g["obs_kind"]) == 0
    if which == "obs_dynamic": ob = np.asarray(g["obs_kind"]) == 1
    if which == "motion": te[:] = True
    for k in ("e6_i", "e6_j", "e6_kind", "e6_meas"): s[k] = s[k][e6]
    for k in ("obs_se3", "obs_point", "obs_kind", "obs_xyz"): s[k] = s[k][ob]
    for k in ("tern_p1", "tern_p2", "tern_h"): s[k] = s[k][te]
    return s


def g2o_chi2(g, n_poses, tmp, save_to=None):
    path = os.path.join(tmp, "graph.g2o")
    g2o_text.write_g2o(path, g, n_poses, precision=17)
    cmd = [os.path.join(ROOT, "oracle", "_ref", "g2o_chi2"), path] + ([save_to] if save_to else [])
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout.split()
    return float(out[0])


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "g2o_ref"])
    cam = synth.SMALL
    sc = synth.Scene(cam=cam, seed=99, flow_noise=0.1, depth_noise=0.01, n_objects=3)
    otr = ol.OracleTracker(ol.track_config(cam, nfeatures=800, max_track_bg=250, max_track_obj=120))
    for k in range(9):
        f = sc.frame(k)
        T, st, rc = otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
        assert rc == 0
    g, n_poses = otr.export_full_graph()
    otr.close()
    se3_opt, pts_opt, its, st = ol.ba_full(g, n_poses)
    g_after = dict(g, se3=np.ascontiguousarray(se3_opt, np.float32).reshape(-1, 16), points=np.ascontiguousarray(pts_opt, np.float32))
    out = {("before_" + k): np.asarray(g[k]) for k in g}
    out["after_se3"], out["after_points"] = g_after["se3"], g_after["points"]
    out["n_poses"] = np.int32(n_poses)
    with tempfile.TemporaryDirectory() as tmp:
        for name, gg in (("before", g), ("after", g_after)):
            out["chi2_" + name] = g2o_chi2(gg, n_poses, tmp)
            prior = g2o_chi2(subset(gg, "prior"), n_poses, tmp)
            out["chi2_" + name + "_prior"] = prior
            for w in SUBSETS[1:]:
                out["chi2_%s_%s" % (name, w)] = g2o_chi2(subset(gg, w), n_poses, tmp) - prior
            print(name, {k: float(v) for k, v in out.items() if k.startswith("chi2_" + name)})
        # a tiny graph printed by g2o itself: 3 poses + 1 motion, a handful of points, every edge type
        keep_pts = np.unique(np.concatenate([g["tern_p1"][:3], g["tern_p2"][:3], g["obs_point"][np.asarray(g["obs_kind"]) == 0][:3]]))
        remap = -np.ones(len(g["points"]), np.int64); remap[keep_pts] = np.arange(len(keep_pts))
        tiny = {k: np.array(g[k], copy=True) for k in g}
        tiny["points"] = g["points"][keep_pts]
        ob = np.isin(g["obs_point"], keep_pts)
        for k in ("obs_se3", "obs_point", "obs_kind", "obs_xyz"): tiny[k] = tiny[k][ob]
        tiny["obs_point"] = remap[tiny["obs_point"]].astype(np.int32)
        for k in ("tern_p1", "tern_p2", "tern_h"): tiny[k] = tiny[k][:3]
        tiny["tern_p1"] = remap[tiny["tern_p1"]].astype(np.int32); tiny["tern_p2"] = remap[tiny["tern_p2"]].astype(np.int32)
        for k in tiny: out["tiny_" + k] = tiny[k]
        out["chi2_tiny"] = g2o_chi2(tiny, n_poses, tmp, save_to=os.path.join(HERE, "g2o_saved_by_reference.g2o"))
    np.savez_compressed(os.path.join(HERE, "g2o_golden.npz"), **out)
    print("sizes: se3", g["se3"].shape, "points", g["points"].shape, "obs", len(g["obs_se3"]), "e6", len(g["e6_i"]), "tern", len(g["tern_p1"]),
          "oracle iterations", its)


if __name__ == "__main__":
    main()
