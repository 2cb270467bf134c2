"""CPU, world_size 2 (gloo): the keyframe-factor all-gather of the multi-sequence configuration leaves every rank with
every rank's flat FullBatch graph, bit for bit."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import fba_synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as ge
    ge._load_pkg()
    from vido_slam_b200 import factors
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, n_poses, _ = fba_synth.make_graph(n_frames=4 + rank, n_objects=1 + rank, seed=10 + rank)
    got, stats = factors.all_gather_factors(g, n_poses)
    ok = len(got) == world and len(stats["bytes_per_rank"]) == world
    for r in range(world):
        gr, npr, _ = fba_synth.make_graph(n_frames=4 + r, n_objects=1 + r, seed=10 + r)
        ok &= got[r][1] == npr
        for k in factors.FBA_KEYS:
            ok &= bool(np.array_equal(np.asarray(got[r][0][k]).reshape(-1), np.asarray(gr[k]).reshape(-1)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_all_gather_factors_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_pack_unpack_roundtrip():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge._load_pkg()
    from vido_slam_b200 import factors
    g, n_poses, _ = fba_synth.make_graph(n_frames=5, n_objects=2, seed=1)
    g2, n2 = factors.unpack_graph(factors.pack_graph(g, n_poses))
    assert n2 == n_poses
    for k in factors.FBA_KEYS:
        assert np.array_equal(g2[k].reshape(-1), np.asarray(g[k]).reshape(-1)) and g2[k].dtype == g[k].dtype
