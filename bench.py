#!/usr/bin/env python
"""bench.py -- frames/s of Tracking + PartialBatchOptimization on the synthetic 1242x375 sequence (BASELINE.json).

A "step" = one chunk of CHUNK consecutive frames of the seeded synthetic sequence pushed through the per-frame driver
(vido_track_frames: gray conversion, ORB, association, init model, pose optimisation, renewal, window BA of every
frame).  The sequence continues across steps, so after the warm-up the sliding window is full (WINDOW_SIZE = 20).

  value : whole-job frames/s with the frames already resident in HBM (device pointers)
  e2e   : the same call with HOST buffers (pinned): H2D copies of image+depth+flow+mask and the D2H of the poses are
          inside the timed region
  roofline : dominant kernel (ba_window_kernel), algorithmic bytes / CUDA-event time against MEASURED_PEAKS.json
  cpu_baseline : the oracle (CPU restatement of the reference path), built ON THIS BOX with the reference's -O3 -march=native,
          single thread like the reference, on bounded samples of the same sequence, with its stage buckets
  vio / dynamic_objects : BASELINE.json configs[2] / configs[3] through the same driver, each with its own resident / end-to-end
          / CPU numbers (measured after the headline arms, outside their timed regions)

  --impl reference : times only the CPU restatement (the reference itself cannot be built here: no OpenCV/Eigen/CSparse
                     C++ in the image) on bounded samples.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

CHUNK = 32   # frames per step (one vido_track_frames call)
BATCH = 16   # frames per front-end batch inside the call (ctx max_batch)
REF_CHUNK = 4
LEG_WARM, LEG_TIMED = 32, 320   # frames of the VIO / dynamic-object legs (warm-up chunk contains the IMU initialisation)
CAM = dict(width=1242, height=375, fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, bf=386.1448)
METRIC = "frames/s Tracking+PartialBA on 1242x375 synth seq"
ORACLE_FLAGS = "-O3 -march=native -ffp-contract=off (oracle/Makefile `native`, built on this box)"


def headline_config():
    """`config` of the bench line, the same object for the GPU arm and the reference arm"""
    bytes_frame = CAM["height"] * CAM["width"] * (3 + 4 + 8 + 4)
    return {"workload": "1242x375 KITTI-shape mono VO, synthetic sequence, ORB + PartialBatchOptimization every frame (BASELINE.json configs[1])",
            "frames_per_step": CHUNK, "front_end_batch": BATCH, "window": 20, "nfeatures": 2500, "max_track_bg": 1000,
            "sequences": "one per GPU (seed 1234+rank)", "l2": f"inputs ({bytes_frame * CHUNK / 1e6:.0f} MB per step) exceed the 126 MB L2",
            "scope": "static scene, VO (the headline); vio / dynamic_objects / full_batch / factor_allgather are measured after the headline arms"}


def load_pkg():
    path = os.path.join(ROOT, "vido-slam_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("vido_slam_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vido_slam_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def native_oracle():
    """the CPU baseline is compiled on the box it is measured on, with the reference's flags (vido_slam/CMakeLists.txt:13-14)"""
    try:
        # named after this host's CPU: a -march=native library built elsewhere (the authoring container's copy travels with the
        # snapshot) must not be picked up here
        import hashlib
        with open("/proc/cpuinfo") as fh:
            cpu = "".join(ln for ln in fh if ln.startswith(("model name", "flags")))[:20000]
        name = "liboracle_native_" + hashlib.sha1(cpu.encode()).hexdigest()[:10] + ".so"
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "native", "NATIVE_OUT=" + name], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        os.environ["VIDO_ORACLE_LIB"] = os.path.join(ROOT, "oracle", name)
        return ORACLE_FLAGS
    except Exception:
        return "-O3 -ffp-contract=off (portable build: `make native` failed on this box)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


STAGES = ("ms_orb", "ms_assoc", "ms_init", "ms_poseopt", "ms_renew", "ms_ba")


def cpu_sample(host_frames, prime, timed, rebuild=1, imu=None, n_objects=0):
    """the CPU restatement over host frames [0, prime + timed): `prime` untimed, `timed` timed.  Returns frames/s and the mean
    stage buckets (feature extraction, association, init model, pose optimisation, map update, local BA -- the reference's
    timing table, src/System.cc:200-233 / src/Tracking.cc:348-359,1121-1139,1319-1330,1451, plus the Frame constructor)"""
    import oracle_lib as ol
    tr = ol.OracleTracker(ol.track_config(CAM, rebuild=rebuild))
    if imu is not None:
        tr.set_imu(imu["Tbc"], imu["noise"])
    k = 0

    def one(k):
        if imu is not None:
            tr.grab_imu(imu["chunks"][k])
            return tr.track(*host_frames(k), timestamp=float(imu["t"][k]))
        return tr.track(*host_frames(k))
    for _ in range(prime):
        one(k); k += 1
    acc = {s: 0.0 for s in STAGES}
    t0 = time.perf_counter()
    for _ in range(timed):
        _, st, _ = one(k); k += 1
        for s in STAGES:
            acc[s] += st[s]
    dt = time.perf_counter() - t0
    tr.close()
    return timed / dt, {s: acc[s] / timed for s in STAGES}


def cv2_front_end_ms(gray):
    """cross-check of the oracle's scalar FAST / resize: OpenCV's own (SIMD) cv::resize pyramid + per-cell cv::FAST loop of
    ORBextractor::ComputeKeyPointsOctTree on one frame, single thread"""
    try:
        import cv2
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import oracle_lib as ol
        from make_orb_golden import cv_level_candidates
        cv2.setNumThreads(1)
        p = ol.default_orb_params()
        H, W = gray.shape
        w, h, _ = ol.level_sizes(W, H, p)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            cur = gray
            for l in range(p.nlevels):
                if l > 0:
                    cur = cv2.resize(cur, (int(w[l]), int(h[l])), interpolation=cv2.INTER_LINEAR)
                cv_level_candidates(cur)
            best = min(best, (time.perf_counter() - t0) * 1e3)
        return {"ms_per_frame": best, "what": "cv2.resize x7 + 1231 cv2 FAST calls from a Python loop (its interpreter overhead included), 1 thread",
                "cv2": cv2.__version__}
    except Exception as e:
        return {"error": str(e)[:120]}


def reference_arm(args, rank, world):
    """CPU restatement of the reference path, single thread like the reference (g2o OpenMP off, Tracking is serial: one
    sequence cannot use more than one core)."""
    if rank != 0:
        return
    flags = native_oracle()
    import synth
    sc = synth.Scene(cam=CAM, seed=1234, flow_noise=0.1, depth_noise=0.01)
    # the window BA reaches its steady-state size (20 poses) after 20 frames: prime at least 24 frames untimed so that
    # the timed sample is the same workload as the GPU arm's (whose warm-up steps cover 96 frames)
    prime = max(args.warmup * REF_CHUNK, 24)
    steps = max(1, min(args.steps, 8))   # bounded sample: <= 32 timed frames (about 2-3 s of CPU work)
    total = prime + steps * REF_CHUNK
    frames = [sc.frame(k) for k in range(total)]
    host = [(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy()) for f in frames]
    t0 = time.perf_counter()
    fps, stage = cpu_sample(lambda k: host[k], prime, steps * REF_CHUNK)
    dt = steps * REF_CHUNK / fps
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
            "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": headline_config(),   # the GPU arm's configuration; a step of this arm is a bounded sample of it (cpu_baseline.sample)
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port", "flags": flags, "stage_ms": stage,
                             "host_cores": os.cpu_count(), "sample_frames_per_step": REF_CHUNK,
                             "sample": f"frames {prime}..{total - 1} of the seed-1234 sequence ({steps} steps, each a {REF_CHUNK}-frame sample of the configuration's {CHUNK}-frame step; CPU restatement; the reference needs OpenCV/Eigen/CSparse C++ which are absent)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bandwidth_kernels():
    """the bandwidth-shaped kernels of the path (front-end, FullBatch linearisation) against the HBM roofline: figures from the
    committed ncu captures (profiles/), not measured by this run -- the line's own `roofline` object is measured live"""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_bandwidth_kernels.json")) as fh:
            return {"roofline_bandwidth_kernels": json.load(fh)}
    except Exception:
        return {}


def descriptor_leg(ctx, torch, dev, gray_host, iters=20):
    """the optional descriptor stage (rBRIEF + Hamming matching; the reference computes no descriptors) beside the
    extraction it builds on: a device-resident batch, extract -> blur + describe -> match every frame against the next one, timed
    per stage with CUDA events on the context stream; the descriptors of frame 0 and its matches are checked against the oracle"""
    import ctypes as C
    import oracle_lib as ol   # the checker of this leg, not the thing measured
    B, H, W = gray_host.shape
    cap = 2564
    d_gray = torch.from_numpy(gray_host).to(dev)
    d_kp = torch.zeros((B, cap, 24), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    out = torch.zeros((3, B, cap), dtype=torch.int32, device=dev)
    vp = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
    stream = torch.cuda.ExternalStream(ctx.stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ms = np.zeros(3)
    torch.cuda.synchronize()
    for it in range(-3, iters):
        ev[0].record(stream)
        ctx.orb_extract_dev(vp(d_gray), B, H * W, W, vp(d_kp), cap, vp(d_n))
        ev[1].record(stream)
        ctx.orb_describe_dev(vp(d_kp), vp(d_n), B, cap, vp(d_desc))
        ev[2].record(stream)
        ctx.hamming_match_dev(vp(d_desc), cap * 32, vp(d_n), vp(d_desc, cap * 32), cap * 32, vp(d_n, 4), B - 1, cap, vp(out[0]), vp(out[1]), vp(out[2]))
        ev[3].record(stream)
        ctx.sync()
        if it >= 0:
            ms += [ev[k].elapsed_time(ev[k + 1]) for k in range(3)]
    ms /= iters
    # the smoothing alone: the same call with all key-point counts zero (the descriptor threads leave at their first test)
    d_zero = torch.zeros(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ms_blur = 0.0
    for it in range(-3, iters):
        ev[0].record(stream)
        ctx.orb_describe_dev(vp(d_kp), vp(d_zero), B, cap, vp(d_desc))
        ev[1].record(stream)
        ctx.sync()
        if it >= 0:
            ms_blur += ev[0].elapsed_time(ev[1])
    ms_blur /= iters
    ctx.orb_describe_dev(vp(d_kp), vp(d_n), B, cap, vp(d_desc), sync=True)   # descriptors back in place for the check below
    n = d_n.cpu().numpy()
    t_c0 = time.perf_counter()
    okp, odesc0 = ol.orb_extract_describe(gray_host[0], ol.default_orb_params())
    t_c1 = time.perf_counter()
    _, odesc1 = ol.orb_extract_describe(gray_host[1], ol.default_orb_params())
    same_desc = bool(n[0] == len(odesc0) and np.array_equal(d_desc[0, :n[0]].cpu().numpy(), odesc0))
    t_c2 = time.perf_counter()
    obi, obd, osd = ol.hamming_match(odesc0, odesc1)
    t_c3 = time.perf_counter()
    got = out[:, 0, :n[0]].cpu().numpy()
    same_match = bool(np.array_equal(got[0], obi) and np.array_equal(got[1], obd) and np.array_equal(got[2], osd))
    w, h, _, _ = ctx.level_info()
    pyr_bytes = int((w.astype(np.int64) * h).sum())
    peak, _ = measured_peak()
    return {"what": "optional descriptor stage beside the extraction: device-resident batch, extract -> 7x7 Gaussian + rBRIEF -> Hamming match of every frame against the next",
            "batch": int(B), "key_points_per_frame": float(n.mean()),
            "ms_per_batch": {"extract": float(ms[0]), "blur_describe": float(ms[1]), "match": float(ms[2])},
            "frames_per_s": {"extract": B / ms[0] * 1e3, "extract_describe": B / (ms[0] + ms[1]) * 1e3, "extract_describe_match": B / ms.sum() * 1e3},
            "blur_describe_algorithmic_GBps": 2 * pyr_bytes * B / (ms[1] * 1e-3) / 1e9, "hbm_peak_GBps": peak,
            "blur_roofline": {"bound": "hbm", "kernel": "blur7_kernel (+ one empty rbrief launch)", "achieved": 2 * pyr_bytes * B / (ms_blur * 1e-3) / 1e9,
                              "peak": peak, "unit": "GB/s", "frac": (2 * pyr_bytes * B / (ms_blur * 1e-3) / 1e9) / peak if peak else None,
                              "avg_launch_ms": float(ms_blur), "algorithmic_bytes": int(2 * pyr_bytes * B)},
            "cpu_port_ms": {"extract_describe_one_frame": (t_c1 - t_c0) * 1e3, "match_one_pair": (t_c3 - t_c2) * 1e3,
                            "note": "the oracle's scalar restatement, one core (not OpenCV's SIMD code)"},
            "descriptors_equal_oracle": same_desc, "matches_equal_oracle": same_match}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample", type=int, default=40, help="frames of the early CPU baseline sample (24 of them untimed)")
    ap.add_argument("--cpu-late", type=int, default=224, help="frame at which the second CPU sample (16 frames) starts; 0 = skip")
    ap.add_argument("--no-legs", action="store_true", help="skip the VIO / dynamic-object / FullBatch legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import synth
    if args.warmup < 3:
        args.warmup = 3
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    pkg = load_pkg()
    ctx = pkg.Context(pkg.default_config(max_batch=BATCH, device=local, **{k: CAM[k] for k in ("width", "height", "fx", "fy", "cx", "cy", "bf")}))

    # ---- synthetic sequence of this rank (one independent sequence per GPU: weak scaling, no data-path collective)
    total = (args.warmup + args.steps) * CHUNK
    legs = rank == 0 and world == 1 and not args.no_legs   # the other configurations are measured at N = 1 only
    nbuf = max(total, LEG_WARM + LEG_TIMED if legs else 0)
    H, W = CAM["height"], CAM["width"]
    img = torch.empty((nbuf, H, W, 3), dtype=torch.uint8, device=dev)
    dep = torch.empty((nbuf, H, W), dtype=torch.float32, device=dev)
    flo = torch.empty((nbuf, H, W, 2), dtype=torch.float32, device=dev)
    msk = torch.zeros((nbuf, H, W), dtype=torch.int32, device=dev)
    h_img, h_dep, h_flo, h_msk = (torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (img, dep, flo, msk))

    def fill(scene, n):
        for k in range(n):
            f = scene.frame(k)
            img[k] = f["gray"].unsqueeze(-1).expand(H, W, 3)
            dep[k] = f["depth_in"]
            flo[k] = f["flow"]
            msk[k] = f["mask"]
        torch.cuda.synchronize()
        for d, s in ((h_img, img), (h_dep, dep), (h_flo, flo), (h_msk, msk)):
            d[:n].copy_(s[:n])
        torch.cuda.synchronize()

    fill(synth.Scene(cam=CAM, seed=1234 + rank, flow_noise=0.1, depth_noise=0.01, device=str(dev)), total)
    bytes_frame = H * W * (3 + 4 + 8 + 4)

    def dev_frames(k0, n, ts=None):
        return [dict(image=img[k].data_ptr(), depth=dep[k].data_ptr(), flow=flo[k].data_ptr(), mask=msk[k].data_ptr(),
                     channels=3, on_device=True, **({"timestamp": float(ts[k])} if ts is not None else {})) for k in range(k0, k0 + n)]

    def host_frames(k0, n, ts=None):
        return [dict(image=h_img[k].numpy(), depth=h_dep[k].numpy(), flow=h_flo[k].numpy(), mask=h_msk[k].numpy(),
                     **({"timestamp": float(ts[k])} if ts is not None else {})) for k in range(k0, k0 + n)]

    def host_gray(k):
        return (h_img[k, :, :, 0].numpy().copy(), h_dep[k].numpy(), h_flo[k].numpy(), h_msk[k].numpy())

    def barrier(collective=True):
        if dist is not None and collective:
            dist.barrier()
        torch.cuda.synchronize()

    def run_arm(make_frames, n_warm, n_steps, ts=None, imu=None, collective=True, before=None):
        """n_warm + n_steps chunks of CHUNK frames through the public API; the copy (host arm) and the front-end of the next chunk
        are announced with vido_track_prefetch and run ahead; everything is inside the timed region"""
        ctx.track_reset()
        if before:
            before()
        k = 0
        last = (n_warm + n_steps) * CHUNK

        # the pointer arrays of every chunk are marshalled once, up front: a C / C++ caller passes its vido_frame_inputs array
        # directly, the Python binding's per-call marshalling (about a millisecond per 32 frames) is not part of the path
        packs = {kk: ctx.pack_frames(make_frames(kk, CHUNK, ts)) for kk in range(0, last, CHUNK)}

        def step(kk, want_stats):
            if kk + 2 * CHUNK <= last:
                ctx.track_prefetch(packs[kk + CHUNK])
            return ctx.track_frames(packs[kk], want_stats=want_stats, imu=(imu[kk:kk + CHUNK] if imu is not None else None))
        for _ in range(n_warm):
            step(k, False); k += CHUNK
        ms0, n0, b0 = ctx.kernel_times()
        l0 = ctx.launches
        barrier(collective)
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.ExternalStream(ctx.stream)
        e0.record(stream)
        stats = []
        for i in range(n_steps):
            # statistics (which make the call wait for its last window solves) only on the last step: the steps before it hand
            # over with the solver queue full, like a caller streaming frames; the closing vido_sync is inside the timed region
            _, st = step(k, i == n_steps - 1); k += CHUNK
            stats += st or []
        ctx.sync()
        e1.record(stream)
        barrier(collective)
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        ms1, n1, b1 = ctx.kernel_times()
        elapsed = max(wall, dev_ms * 1e-3)
        if dist is not None and collective:
            t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed = float(t.item())
        return elapsed, ms1 - ms0, n1 - n0, b1 - b0, ctx.launches - l0, stats

    sampler = ClockSampler(local)
    sampler.start()
    el_dev, kms, kn, kbytes, launches, stats = run_arm(dev_frames, args.warmup, args.steps)
    el_e2e, _, _, _, _, _ = run_arm(host_frames, args.warmup, args.steps)
    sampler.stop_flag = True

    # ---- outside the timed region: the once-per-sequence stages --------------------------------------------------
    # (1) keyframe factors of this rank's sequence (flat FullBatch graph); with N > 1 ONE all-gather (NCCL) leaves the
    #     factors of all sequences on every rank (BASELINE.json configs[4]); (2) FullBatchOptimization of the own block
    extra = {}
    try:
        g, npo = ctx.export_full_graph()
        if dist is not None:
            from vido_slam_b200 import factors
            gathered, fst = factors.all_gather_factors(g, npo, device=dev)
            extra["factor_allgather"] = {"ranks": len(gathered), "bytes_per_rank": fst["bytes_per_rank"],
                                         "allgather_ms": fst["allgather_ms"],
                                         "total_poses": int(sum(n for _, n in gathered))}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st_fb, sizes = ctx.full_batch()
        dt_fb = time.perf_counter() - t0
        # the same graph once more through vido_ba_full (the exported copy still holds the initial values): the first solve of a
        # context also pays the device allocation of its arena, whose cost varies from a millisecond to hundreds (driver-side)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, st_w = ctx.ba_full(g, npo)
        dt_fb_warm = time.perf_counter() - t0
        if dist is not None:   # what the gathered factors are for: solve the neighbour's block, compare with the neighbour's own result
            diff, its = factors.cross_check(ctx, gathered, rank, ctx.map_poses_rf(), device=dev)
            extra["factor_allgather"]["neighbour_block_max_abs_diff"] = diff
            extra["factor_allgather"]["neighbour_block_iterations"] = its
        rec = st_fb.records()
        extra["full_batch"] = {"frames": int(sizes[0]), "points": int(sizes[2]), "observations": int(sizes[3]),
                               "iterations": int(st_fb.iterations), "trials": int(st_fb.total_trials), "ms": dt_fb * 1e3, "ms_same_graph_again": dt_fb_warm * 1e3, "iterations_again": int(st_w.iterations),
                               "chi2_first": rec[0][0] if rec else None, "chi2_last": rec[-1][0] if rec else None}
        if legs:   # the same graph through the CPU restatement of FullBatchOptimization
            import oracle_lib as ol
            native_oracle()
            t0 = time.perf_counter()
            if int(sizes[3]) <= 600000:   # bounded: the CPU solve of larger graphs takes minutes
                _, _, its_cpu, _ = ol.ba_full(g, npo)
                extra["full_batch"]["cpu_port_ms"] = (time.perf_counter() - t0) * 1e3
                extra["full_batch"]["cpu_port_iterations"] = int(its_cpu)
    except Exception as e:  # reported, never fatal for the headline metric
        extra.setdefault("full_batch", {})["error"] = str(e)[:200]

    # ---- the reference's own call pattern: one frame per call, window optimisation finished before the call returns (what the
    #      C++ facade System::TrackRGBD does, host buffers in, pose out) -- latency-bound, no look-ahead, no pipelining across frames
    frame_at_a_time = None
    if world == 1 and legs:
        try:
            n_w, n_t = 32, 96
            singles = [ctx.pack_frames(host_frames(k, 1)) for k in range(n_w + n_t)]
            ctx.track_reset()
            for k in range(n_w):
                ctx.track_frames(singles[k], want_stats=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(n_w, n_w + n_t):
                ctx.track_frames(singles[k], want_stats=True)
            dt1 = time.perf_counter() - t0
            frame_at_a_time = {"value": n_t / dt1, "unit": "frames/s", "frames_timed": n_t, "ms_per_frame": 1e3 * dt1 / n_t,
                               "what": "one vido_track_frames call per frame with statistics (= System::TrackRGBD of the C++ facade): host buffers in, "
                                       "front-end, tracking and the frame's window optimisation all finished when the call returns"}
        except Exception as e:
            frame_at_a_time = {"error": str(e)[:200]}

    flags = native_oracle() if rank == 0 else None

    def leg(name, scene, workload, ts=None, imu=None, imu_cpu=None, before=None):
        """one of BASELINE.json's other single-GPU configurations through the same driver: resident, end-to-end and CPU numbers"""
        out = {"unit": "frames/s", "workload": workload, "frames_timed": LEG_TIMED}
        try:
            fill(scene, LEG_WARM + LEG_TIMED)
            nw, ns = LEG_WARM // CHUNK, LEG_TIMED // CHUNK
            el, _, _, _, _, st = run_arm(dev_frames, nw, ns, ts=ts, imu=imu, collective=False, before=before)
            out["value"] = LEG_TIMED / el
            out["host_ms_per_frame"] = {k: float(np.mean([x[k] for x in st])) for k in ("ms_init", "ms_poseopt", "ms_renew", "ms_ba")}
            out["objects_per_frame"] = float(np.mean([x["n_objects_ok"] for x in st]))
            out["object_features_per_frame"] = float(np.mean([x["n_dyn_features"] for x in st]))
            el2, _, _, _, _, _ = run_arm(host_frames, nw, ns, ts=ts, imu=imu, collective=False, before=before)
            out["e2e"] = {"value": LEG_TIMED / el2, "unit": "frames/s", "h2d_bytes_per_step": bytes_frame * CHUNK, "d2h_bytes_per_step": 64 * CHUNK}
            fps, stage = cpu_sample(host_gray, LEG_WARM, 16, imu=imu_cpu)
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port", "flags": flags, "stage_ms": stage,
                                   "sample": f"frames {LEG_WARM}..{LEG_WARM + 15} of the same sequence"}
            out["vs_cpu"] = {"resident": out["value"] / fps, "e2e": out["e2e"]["value"] / fps}
        except Exception as e:
            out["error"] = str(e)[:200]
        return out

    if legs:
        # (2b) VIO mode (BASELINE.json configs[2]): 10 frames/s, 200 Hz IMU riding on the camera; preintegration of every
        #      frame, InitializeIMU after frame 21 (inside the warm-up chunk), then tracking in the gravity-aligned metric map
        try:
            import imu_synth
            nv = LEG_WARM + LEG_TIMED
            smp, ft, Tbc, _ = imu_synth.make_vio_sequence(nv, fps=10.0, pose_np=imu_synth.vio_camera_pose_10fps_np)
            chunks = imu_synth.imu_chunks(smp, ft)
            extra["vio"] = leg("vio", synth.Scene(cam=CAM, seed=777, device=str(dev), pose_fn=imu_synth.vio_camera_pose_10fps),
                               "VIO mode (BASELINE.json configs[2]): noise-free depth / flow, 20 IMU samples per frame (200 Hz at 10 frames/s); timed after the IMU initialisation",
                               ts=ft, imu=chunks, imu_cpu={"Tbc": Tbc, "noise": imu_synth.NOISE, "chunks": chunks, "t": ft},
                               before=lambda: ctx.track_set_imu(Tbc, imu_synth.NOISE))
            ist = ctx.imu_state()
            extra["vio"].update({"imu_hz": 200, "imu_initialized": int(ist.initialized), "init_frame": int(ist.init_frame), "scale": float(ist.scale),
                                 "gyro_bias": [float(x) for x in ist.bg], "inertial_lm": [int(ist.lm_iterations), int(ist.lm_trials)]})
        except Exception as e:
            extra["vio"] = {"error": str(e)[:200]}
        finally:
            try:
                ctx.track_reset()
                ctx.track_set_imu(None, None)   # back to sensor = RGBD for the remaining legs
            except Exception:
                pass
        # (3) the same driver on a scene with 5 moving objects (BASELINE.json configs[3])
        extra["dynamic_objects"] = leg("dynamic_objects", synth.Scene(cam=CAM, seed=4321, flow_noise=0.1, depth_noise=0.01, device=str(dev), n_objects=5),
                                       "5 moving objects / frame, masks + flow, joint static / object-motion tracking (BASELINE.json configs[3])")
        try:   # joint FullBatch (camera poses, static points, object points, object motions) over the dynamic frames just tracked
            ctx.track_reset()
            ctx.track_frames(dev_frames(0, 3 * CHUNK), want_stats=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st_d, sz_d = ctx.full_batch()
            dt_d = time.perf_counter() - t0
            recd = st_d.records()
            extra["dynamic_objects"]["full_batch"] = {"frames": int(sz_d[0]), "object_motions": int(sz_d[1]), "points": int(sz_d[2]),
                                                      "observations": int(sz_d[3]), "ternary_edges": int(sz_d[5]),
                                                      "iterations": int(st_d.iterations), "trials": int(st_d.total_trials),
                                                      "cg_iterations": int(st_d.pad), "ms": dt_d * 1e3,
                                                      "chi2_first": recd[0][0] if recd else None, "chi2_last": recd[-1][0] if recd else None}
        except Exception as e:
            extra["dynamic_objects"]["full_batch"] = {"error": str(e)[:200]}
    frames_total = args.steps * CHUNK * world
    value = frames_total / el_dev
    e2e = frames_total / el_e2e

    if rank == 0:
        peak, peak_kind = measured_peak()
        ba_ms = kms[3] / max(kn[3], 1)
        achieved = (kbytes / max(kn[3], 1)) / (ba_ms * 1e-3) / 1e9 if ba_ms > 0 else 0.0
        share = {n: float(v) for n, v in zip(("orb_front_end", "init_model", "pose_opt", "window_ba"), kms)}
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), if any
        traffic = None
        for name in ("r2_ba_traffic.json", "r1_ba_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as fh:
                    traffic = float(json.load(fh)["dram_bytes_per_launch"])
                break
            except Exception:
                pass
        yardstick = None   # what bounds the dominant kernel instead of HBM: committed probe figures (profiles/), not measured by this run
        try:
            with open(os.path.join(ROOT, "profiles", "r2_ba_yardstick.json")) as fh:
                yardstick = json.load(fh)
        except Exception:
            pass
        # CPU baseline: the headline sequence again (the legs reused the buffers)
        fill(synth.Scene(cam=CAM, seed=1234, flow_noise=0.1, depth_noise=0.01, device=str(dev)), min(total, max(args.cpu_sample, args.cpu_late + 16)))
        ncpu = min(args.cpu_sample, total)
        skip = min(24, ncpu // 2)
        cpu_fps, cpu_stage = cpu_sample(host_gray, skip, ncpu - skip, rebuild=1)
        # second number "for honesty" (SURVEY 8d): the same CPU path with incremental tracklet bookkeeping instead of the
        # reference's rebuild from frame 0 every frame (only matters for long sequences)
        cpu_fps_inc, _ = cpu_sample(host_gray, skip, ncpu - skip, rebuild=0)
        late = None
        if args.cpu_late and args.cpu_late + 16 <= total:   # the rebuild grows with the frame index: a second, later sample
            fa, sa = cpu_sample(host_gray, args.cpu_late, 16, rebuild=1)
            late = {"first_frame": args.cpu_late, "frames": 16, "value_rebuild": fa, "stage_ms_rebuild": sa}
        cv2_ms = cv2_front_end_ms(h_img[skip, :, :, 0].numpy().copy())
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": el_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": headline_config(),
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": bytes_frame * CHUNK, "d2h_bytes_per_step": 64 * CHUNK},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "ba_window_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_kind,
                         "avg_launch_ms": ba_ms, "device_ms_by_stage": share, "latency_yardstick": yardstick,
                         "note": "the window problem is shared-memory / L2 resident: the kernel is bound by FP64 issue and dependent-latency chains, not by HBM (DESIGN.md section 6)"},
            "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": 1, "kind": "port", "flags": flags, "stage_ms": cpu_stage,
                             "sample": f"frames {skip}..{ncpu - 1} of the same sequence through the CPU restatement (single thread, like the reference)",
                             "value_incremental_tracklets": cpu_fps_inc, "late_sample": late, "host_cores": os.cpu_count(),
                             "one_sequence_per_core_estimate": cpu_fps * (os.cpu_count() or 1),
                             "cv2_front_end_cross_check": cv2_ms},
            **({"frame_at_a_time": frame_at_a_time} if frame_at_a_time else {}),
            **bandwidth_kernels(),
            **extra,
            "host_ms_per_frame": {k: float(np.mean([x[k] for x in stats])) for k in ("ms_orb", "ms_init", "ms_poseopt", "ms_renew", "ms_ba")},
            "ba_per_frame": {"iterations": float(np.mean([s["ba_iterations"] for s in stats])), "obs": float(np.mean([s["ba_obs"] for s in stats])),
                             "points": float(np.mean([s["ba_points"] for s in stats]))},
        }
        if legs:   # last: nothing of the headline depends on it
            try:
                line["descriptors"] = descriptor_leg(ctx, torch, dev, np.ascontiguousarray(h_img[:BATCH, :, :, 0].numpy()))
            except Exception as e:
                line["descriptors"] = {"error": str(e)[:200]}
            try:   # the same at a batch that fills the machine (a second context; the HBM-shaped kernel of the stage against its roofline)
                nb = min(64, total)
                ctx64 = pkg.Context(pkg.default_config(max_batch=nb, device=local, **{k: CAM[k] for k in ("width", "height", "fx", "fy", "cx", "cy", "bf")}))
                line["descriptors_batch64"] = descriptor_leg(ctx64, torch, dev, np.ascontiguousarray(h_img[:nb, :, :, 0].numpy()))
                ctx64.close()
            except Exception as e:
                line["descriptors_batch64"] = {"error": str(e)[:200]}
            try:   # the opt-in sliding-window variant of the smoothing kernel (VIDO_BLUR=v2; CPU-emulated only when it was committed):
                   # measured and checked against the oracle here so that the round's record holds both numbers; last on purpose
                os.environ["VIDO_BLUR"] = "v2"
                ctx64 = pkg.Context(pkg.default_config(max_batch=nb, device=local, **{k: CAM[k] for k in ("width", "height", "fx", "fy", "cx", "cy", "bf")}))
                line["descriptors_batch64_blur_v2"] = descriptor_leg(ctx64, torch, dev, np.ascontiguousarray(h_img[:nb, :, :, 0].numpy()))
                ctx64.close()
            except Exception as e:
                line["descriptors_batch64_blur_v2"] = {"error": str(e)[:200]}
            finally:
                os.environ.pop("VIDO_BLUR", None)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
