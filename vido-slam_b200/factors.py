"""Keyframe-factor exchange for the multi-sequence configuration (BASELINE.json configs[4]).

One process per GPU tracks one independent sequence; there is no collective on the per-frame path.  After tracking, every
rank packs the flat FullBatch graph of its Map (vido_fba_problem, include/vido_b200.h) into one byte buffer and ONE
all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests) leaves all graphs on all ranks.  The joint system is
block-diagonal by sequence, so every rank then solves its own block with vido_ba_full: the result equals independent
FullBatchOptimization runs (the reference has no cross-sequence coupling at all).

torch.distributed is plumbing here; the product is the graph layout and the solver behind the C-ABI.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import FBA_F32, FBA_KEYS, FBA_WIDTH

_HDR = 16  # int64 words: magic, n_poses, then the element count of every FBA_KEYS array


def pack_graph(g, n_poses):
    """dict keyed by FBA_KEYS -> 1-D uint8 numpy buffer (header + arrays, each 8-byte aligned)"""
    hdr = np.zeros(_HDR, np.int64)
    hdr[0], hdr[1] = 0x5649444F, n_poses
    parts = []
    for i, k in enumerate(FBA_KEYS):
        a = np.ascontiguousarray(g[k], np.float32 if k in FBA_F32 else np.int32).reshape(-1)
        hdr[2 + i] = a.size
        b = a.view(np.uint8)
        pad = (-b.size) % 8
        parts.append(b)
        if pad:
            parts.append(np.zeros(pad, np.uint8))
    return np.concatenate([hdr.view(np.uint8)] + parts)


def unpack_graph(buf):
    """inverse of pack_graph: (dict keyed by FBA_KEYS, n_poses)"""
    buf = np.ascontiguousarray(buf, np.uint8)
    hdr = buf[:_HDR * 8].view(np.int64)
    if hdr[0] != 0x5649444F:
        raise ValueError("not a keyframe-factor buffer")
    off, g = _HDR * 8, {}
    for i, k in enumerate(FBA_KEYS):
        n = int(hdr[2 + i])
        a = buf[off:off + 4 * n].view(np.float32 if k in FBA_F32 else np.int32).copy()
        g[k] = a.reshape(-1, FBA_WIDTH[k]) if k in FBA_WIDTH else a
        off += 4 * n + ((-4 * n) % 8)
    return g, int(hdr[1])


def all_gather_factors(g, n_poses, device=None, group=None):
    """every rank contributes its graph; returns ([(graph, n_poses) for every rank], stats).  The payload moves in ONE
    all-gather of equal-size byte buffers (padded to the largest rank, whose size comes from a 1-word metadata exchange)."""
    world = dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    mine = torch.from_numpy(pack_graph(g, n_poses))
    size = torch.tensor([mine.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = (max(sizes) + 255) // 256 * 256
    send = torch.zeros(cap, dtype=torch.uint8, device=dev)
    send[:mine.numel()] = mine.to(dev)
    recv = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    dist.all_gather_into_tensor(recv, send, group=group)
    ms = None
    if dev.type == "cuda":
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
    host = recv.cpu().numpy()
    out = [unpack_graph(host[r * cap:r * cap + sizes[r]]) for r in range(world)]
    return out, dict(bytes_per_rank=sizes, padded_bytes=cap, allgather_ms=ms)


def cross_check(ctx, gathered, rank, own_refined, device=None, group=None):
    """What the gathered factors are for: rank r solves the block of rank (r + 1) % N from the GATHERED graph with the flat-graph
    solver (vido_ba_full) and compares the refined camera poses with the ones that rank obtained from its own Map
    (vido_full_batch).  The refined poses of every rank travel in a second, small all-gather.  Returns the largest absolute
    difference this rank saw (the joint system is block-diagonal by sequence: it must be solver round-off)."""
    world = dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    counts = [int(n) for _, n in gathered]
    cap = max(counts)
    mine = torch.zeros(cap * 16, dtype=torch.float32, device=dev)
    own = torch.from_numpy(np.ascontiguousarray(own_refined, np.float32).reshape(-1))
    mine[:own.numel()] = own.to(dev)
    allp = torch.empty(world * cap * 16, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(allp, mine, group=group)
    nb = (rank + 1) % world
    g, npo = gathered[nb]
    se3, _, st = ctx.ba_full(g, npo)
    want = allp.cpu().numpy().reshape(world, cap, 16)[nb, :npo]
    got = se3[:npo].reshape(npo, 16)
    return float(np.abs(got - want).max()), int(st.iterations)
