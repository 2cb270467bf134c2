"""GPU: vido_ba_full / vido_full_batch (Optimizer::FullBatchOptimization, src/Optimizer.cc:1235-2178) against the oracle:
same LM trajectory (iterations, trials, robust chi2 per iteration), estimates within 1e-4 relative."""
import numpy as np
import pytest

import fba_synth
import oracle_lib as ol
import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4
CAM = synth.KITTI


def _same_lm(st_gpu, st_ref, rel=1e-6):
    a, b = st_gpu.records(), st_ref.records()
    assert st_gpu.iterations == st_ref.iterations and st_gpu.total_trials == st_ref.total_trials, (st_gpu.iterations, st_ref.iterations)
    assert len(a) == len(b)
    for (c1, l1, t1), (c2, l2, t2) in zip(a, b):
        assert t1 == t2
        assert abs(c1 - c2) <= rel * max(abs(c2), 1e-12) and abs(l1 - l2) <= 1e-6 * abs(l2)


@pytest.mark.parametrize("n_frames,n_objects,seed,n_static", [(6, 2, 3, 60), (5, 0, 5, 60), (9, 3, 11, 60), (36, 0, 7, 400)])
def test_flat_graph_matches_oracle(pkg, n_frames, n_objects, seed, n_static):
    # (the 36-frame case is banded: tracks span 4 frames, so the factorisation skips everything outside the envelope)
    g, n_poses, truth = fba_synth.make_graph(n_frames=n_frames, n_objects=n_objects, seed=seed, n_static=n_static)
    se3_ref, pts_ref, its, st_ref = ol.ba_full(g, n_poses)
    ctx = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    se3, pts, st = ctx.ba_full(g, n_poses)
    _same_lm(st, st_ref)
    assert np.abs(se3 - se3_ref).max() <= REL_TOL * max(np.abs(se3_ref).max(), 1.0)
    assert np.abs(pts - pts_ref).max() <= REL_TOL * max(np.abs(pts_ref).max(), 1.0)
    assert np.abs(se3[:n_poses, :3, 3] - truth["Twc"][:, :3, 3]).max() < 0.05      # and right in absolute terms
    ctx.close()


@pytest.mark.parametrize("n_frames,n_objects,seed", [(6, 2, 3), (9, 3, 11)])
def test_matrix_free_solver_matches_oracle(pkg, n_frames, n_objects, seed):
    """solver = 2: implicit Schur complement + preconditioned CG instead of the explicit reduced system; the linear systems are
    solved to 1e-13, so the LM trajectory is the oracle's as well"""
    g, n_poses, truth = fba_synth.make_graph(n_frames=n_frames, n_objects=n_objects, seed=seed)
    se3_ref, pts_ref, its, st_ref = ol.ba_full(g, n_poses)
    ctx = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    se3, pts, st = ctx.ba_full(g, n_poses, solver=2)
    assert st.pad > 0                                  # CG iterations were spent
    _same_lm(st, st_ref, rel=1e-5)
    assert np.abs(se3 - se3_ref).max() <= REL_TOL * max(np.abs(se3_ref).max(), 1.0)
    assert np.abs(pts - pts_ref).max() <= REL_TOL * max(np.abs(pts_ref).max(), 1.0)
    ctx.close()


def test_bad_graphs_are_rejected(pkg):
    g, n_poses, _ = fba_synth.make_graph(n_frames=4, n_objects=1, seed=2)
    ctx = pkg.Context(pkg.default_config(width=640, height=480, max_batch=1))
    bad = dict(g); bad["obs_point"] = g["obs_point"].copy(); bad["obs_point"][0] = 10 ** 6
    with pytest.raises(pkg.VidoError):
        ctx.ba_full(bad, n_poses)
    bad = dict(g); bad["tern_p1"] = g["tern_p1"].copy(); bad["tern_p1"][1] = g["tern_p1"][0]   # two successors of one point
    with pytest.raises(pkg.VidoError):
        ctx.ba_full(bad, n_poses)
    ctx.close()


def test_pipeline_full_batch_matches_oracle(pkg):
    """track a dynamic sequence, then FullBatch on the Map: same flat graph as the oracle's tracker, same solution"""
    n = 6
    sc = synth.Scene(cam=CAM, seed=1234, flow_noise=0.1, depth_noise=0.01, n_objects=5)
    frames = [sc.frame(k) for k in range(n)]
    otr = ol.OracleTracker(ol.track_config(CAM))
    for f in frames:
        otr.track(f["gray"].numpy(), f["depth_in"].numpy(), f["flow"].numpy(), f["mask"].numpy())
    ctx = pkg.Context(pkg.default_config(width=CAM["width"], height=CAM["height"], fx=CAM["fx"], fy=CAM["fy"], cx=CAM["cx"],
                                         cy=CAM["cy"], bf=CAM["bf"], max_batch=3))
    ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(),
                           mask=f["mask"].numpy().copy()) for f in frames], want_stats=False)
    # (1) the keyframe-factor graph: identical structure, float payload within the tracking tolerance
    g0, np0 = otr.export_full_graph()
    g1, np1 = ctx.export_full_graph()
    assert np0 == np1 == n
    for k in ol.FBA_KEYS:
        assert g0[k].shape == g1[k].shape, k
        if g0[k].dtype == np.int32:
            assert np.array_equal(g0[k], g1[k]), k
        elif g0[k].size:
            assert np.abs(g0[k] - g1[k]).max() <= REL_TOL * max(np.abs(g0[k]).max(), 1.0), k
    assert len(g0["tern_p1"]) > 1000 and len(g0["obs_se3"]) > 5000
    # (2) the solver on the SAME graph (the oracle's), bounded to a few iterations to keep the CPU side short
    se3_ref, pts_ref, its, st_ref = ol.ba_full(g0, np0, max_iterations=6)
    se3, pts, st = ctx.ba_full(g0, np0, max_iterations=6)
    _same_lm(st, st_ref)
    assert np.abs(se3 - se3_ref).max() <= REL_TOL * max(np.abs(se3_ref).max(), 1.0)
    assert np.abs(pts - pts_ref).max() <= REL_TOL * max(np.abs(pts_ref).max(), 1.0)
    se3_cg, pts_cg, st_cg = ctx.ba_full(g0, np0, max_iterations=6, solver=2)
    assert st_cg.iterations == st_ref.iterations and st_cg.total_trials == st_ref.total_trials
    assert np.abs(se3_cg - se3_ref).max() <= REL_TOL * max(np.abs(se3_ref).max(), 1.0)
    assert np.abs(pts_cg - pts_ref).max() <= REL_TOL * max(np.abs(pts_ref).max(), 1.0)
    # (3) vido_full_batch on the context's own Map runs to the reference's stop rule and improves the robust chi2
    st_full, sizes = ctx.full_batch()
    rec = st_full.records()
    assert list(sizes) == [np0, g0["se3"].shape[0] - np0, g0["points"].shape[0], len(g0["obs_se3"]), len(g0["e6_i"]), len(g0["tern_p1"])]
    assert st_full.iterations >= 5 and rec[-1][0] < rec[0][0]
    P, Prf = ctx.map_poses(), ctx.map_poses_rf()
    assert np.array_equal(P[0], Prf[0]) and np.isfinite(Prf).all() and np.abs(P - Prf).max() < 0.2
    for k in range(1, n):
        assert ctx.map_objects_rf(k).shape == ctx.map_objects(k)[2].shape
    otr.close(); ctx.close()


def _gathered_worker(rank, world, port, q):
    import os
    import sys
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    import __graft_entry__ as ge
    pkg = ge._load_pkg()
    from vido_slam_b200 import factors
    import synth
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cam = synth.SMALL
    sc = synth.Scene(cam=cam, seed=300 + rank, flow_noise=0.1, depth_noise=0.01, n_objects=rank)   # rank 1 has a moving object
    frames = [sc.frame(k) for k in range(8 + 2 * rank)]
    ctx = pkg.Context(pkg.default_config(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"],
                                         bf=cam["bf"], max_batch=4))
    ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(), mask=f["mask"].numpy()) for f in frames],
                     want_stats=False)
    g, npo = ctx.export_full_graph()
    gathered, _ = factors.all_gather_factors(g, npo)
    ctx.full_batch()
    diff, its = factors.cross_check(ctx, gathered, rank, ctx.map_poses_rf())
    q.put((rank, diff, its, len(gathered)))
    ctx.close()
    dist.destroy_process_group()


def test_gathered_block_solves_like_its_owner():
    """BASELINE.json configs[4]: after the keyframe-factor all-gather every rank holds every sequence's graph; rank r solves the
    GATHERED block of rank r+1 and must reproduce that rank's own FullBatchOptimization (two processes share cuda:0 here, the
    exchange runs over gloo; bench.py --gpus N does the same over NCCL)"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mctx = mp.get_context("spawn")
    q = mctx.Queue()
    procs = [mctx.Process(target=_gathered_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, diff, its, n in res:
        assert n == 2 and its >= 1
        assert diff <= 1e-4, (rank, diff)


def test_full_batch_writes_the_reference_graph_files(pkg, tmp_path, monkeypatch):
    """VIDO_SAVE_G2O=1: vido_full_batch writes dynamic_slam_graph_before_opt.g2o / ..._after_opt.g2o into the working directory like
    Optimizer::FullBatchOptimization does (src/Optimizer.cc:1937,1939); the 'after' file holds the refined camera poses"""
    import importlib
    import synth
    g2o_text = importlib.import_module("vido-slam_b200.g2o_text")
    cam, n = synth.SMALL, 8
    sc = synth.Scene(cam=cam, seed=5, flow_noise=0.1, depth_noise=0.01, n_objects=2)
    frames = [sc.frame(k) for k in range(n)]
    ctx = pkg.Context(pkg.default_config(width=cam["width"], height=cam["height"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"],
                                         bf=cam["bf"], max_batch=4))
    ctx.track_frames([dict(image=f["gray"].numpy(), depth=f["depth_in"].numpy(), flow=f["flow"].numpy(), mask=f["mask"].numpy().copy())
                      for f in frames], want_stats=False)
    g, n_poses = ctx.export_full_graph()
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("VIDO_SAVE_G2O", "1")
    st, sizes = ctx.full_batch()
    before = g2o_text.read_g2o(tmp_path / "dynamic_slam_graph_before_opt.g2o")
    after = g2o_text.read_g2o(tmp_path / "dynamic_slam_graph_after_opt.g2o")
    assert len(before["se3"]) == len(g["se3"]) == sizes[0] + sizes[1] and len(before["points"]) == len(g["points"])
    assert len(before["obs"]) == len(g["obs_se3"]) and len(before["tern"]) == len(g["tern_p1"]) and len(before["prior"]) == 1
    T0 = np.asarray(g["se3"], np.float64).reshape(-1, 4, 4)
    for k in (0, n_poses - 1, len(T0) - 1):
        assert np.abs(before["se3"][g2o_text.FIRST_ID + k] - T0[k]).max() < 2e-5 * max(np.abs(T0[k]).max(), 1.0)   # 6 printed digits
    P = ctx.map_poses_rf()
    for k in (1, n_poses - 1):
        assert np.abs(after["se3"][g2o_text.FIRST_ID + k] - P[k]).max() < 2e-5 * max(np.abs(P[k]).max(), 1.0)
    ctx.close()
