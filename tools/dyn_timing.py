"""debug: host-path stage times of the 5-object configuration (VIDO_HOST_TIMING=1 python tools/dyn_timing.py)"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
pkg = importlib.import_module("vido-slam_b200")
cam = synth.KITTI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
sc = synth.Scene(cam=cam, seed=1234, flow_noise=0.1, depth_noise=0.01, n_objects=5, device="cuda")
frames = []
for k in range(n):
    f = sc.frame(k)
    frames.append(dict(image=f["gray"].cpu().numpy(), depth=f["depth_in"].cpu().numpy(), flow=f["flow"].cpu().numpy(), mask=f["mask"].cpu().numpy()))
ctx = pkg.Context(pkg.default_config(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"], bf=cam["bf"], max_batch=16))
ctx.track_frames(frames[:32])
t0 = time.perf_counter()
T, st = ctx.track_frames(frames[32:])
el = time.perf_counter() - t0
print("dynamic frames/s", (n - 32) / el, "objects", np.mean([s["n_objects"] for s in st]), "dyn feats", np.mean([s["n_dyn_features"] for s in st]))
