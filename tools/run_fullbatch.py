"""Small driver for profiling the full-sequence optimisation under ncu: tracks N frames of the synthetic scene (optionally with
moving objects) and runs vido_full_batch once.  usage: python tools/run_fullbatch.py [frames] [objects]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as ge
import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nobj = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pkg = ge._load_pkg()
CAM = synth.KITTI
sc = synth.Scene(cam=CAM, seed=1234, flow_noise=0.1, depth_noise=0.01, n_objects=nobj, device="cuda")
ctx = pkg.Context(pkg.default_config(max_batch=8, **{k: CAM[k] for k in ("width", "height", "fx", "fy", "cx", "cy", "bf")}))
frames = [sc.frame(k) for k in range(n)]
host = [dict(image=f["gray"].cpu().numpy(), depth=f["depth_in"].cpu().numpy(), flow=f["flow"].cpu().numpy(), mask=f["mask"].cpu().numpy())
        for f in frames]
t0 = time.perf_counter()
ctx.track_frames(host, want_stats=False)
t1 = time.perf_counter()
st, sizes = ctx.full_batch()
t2 = time.perf_counter()
rec = st.records()
print(f"tracked {n} frames in {1e3 * (t1 - t0):.1f} ms; full batch: sizes {list(sizes)} iterations {st.iterations} trials {st.total_trials} "
      f"{1e3 * (t2 - t1):.1f} ms chi2 {rec[0][0]:.4f} -> {rec[-1][0]:.4f}  (CG iterations of the matrix-free path: {st.pad})")
ctx.close()
