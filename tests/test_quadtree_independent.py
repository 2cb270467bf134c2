"""CPU: the quad-tree culling of the ORB extractor (ORBextractor::DistributeOctTree + ExtractorNode::DivideNode,
/root/reference/vido_slam/src/ORBextractor.cc:471-753) restated a second time, independently of oracle/orb_oracle.cc, in plain
Python, and compared with the oracle's whole extraction level by level (key point list, order, responses).

Tie-break: the reference sorts (count, node pointer) pairs, i.e. by heap address among equal counts -- unspecified behaviour.
Both restatements use the rule documented in oracle/orb_oracle.cc: among nodes of equal count the one created LATER is
expanded first.  Everything else (the roots, the split geometry with ceil(width / 2), children pushed to the FRONT of the node
list, the parent erased, "first maximum response wins" per node, the final order = list order) follows the cited lines."""
import math

import numpy as np
import pytest

import oracle_lib as ol

EDGE = 19


def f32(v):
    return float(np.float32(v))


def split(node, cand):
    """DivideNode (:471-527): integer corners, halves rounded up; a key goes left/up when strictly below the middle"""
    ulx, uly, urx, bry = node["ulx"], node["uly"], node["urx"], node["bry"]
    hx = int(math.ceil(f32(urx - ulx) / 2)); hy = int(math.ceil(f32(bry - uly) / 2))
    mx, my = ulx + hx, uly + hy
    kids = [dict(ulx=ulx, uly=uly, urx=mx, bry=my, keys=[]), dict(ulx=mx, uly=uly, urx=urx, bry=my, keys=[]),
            dict(ulx=ulx, uly=my, urx=mx, bry=bry, keys=[]), dict(ulx=mx, uly=my, urx=urx, bry=bry, keys=[])]
    for k in node["keys"]:
        x, y = cand[k][0], cand[k][1]
        kids[(0 if x < mx else 1) + (0 if y < my else 2)]["keys"].append(k)
    return kids


def distribute(cand, minX, maxX, minY, maxY, N):
    """cand: list of (x, y, response) in detection order; returns the indices kept, in the reference's output order"""
    if not cand:
        return []
    nIni = int(round(f32(maxX - minX) / (maxY - minY)))      # C round(): half away from zero; same as Python here (no .5 ties hit)
    hX = np.float32(maxX - minX) / np.float32(nIni)
    serial = [0]

    def stamp(n):
        n["serial"] = serial[0]; serial[0] += 1
        n["done"] = len(n["keys"]) == 1
        return n
    nodes = [stamp(dict(ulx=int(hX * np.float32(i)), uly=0, urx=int(hX * np.float32(i + 1)), bry=maxY - minY, keys=[])) for i in range(nIni)]
    for k, c in enumerate(cand):
        nodes[min(int(np.float32(c[0]) / hX), nIni - 1)]["keys"].append(k)
    nodes = [n for n in nodes if n["keys"]]
    for n in nodes:
        n["done"] = len(n["keys"]) == 1

    def expand(n, nodes, pending):
        """children with keys go to the FRONT one after the other (so the last child ends up first); the parent is removed"""
        for kid in split(n, cand):
            if kid["keys"]:
                stamp(kid)
                nodes.insert(0, kid)
                if len(kid["keys"]) > 1:
                    pending.append(kid)
        nodes.remove(n)

    while True:
        prev = len(nodes)
        pending = []
        for n in [m for m in nodes if not m["done"]]:      # the list as it is when the pass starts: new children sit in front
            expand(n, nodes, pending)
        n_expand = len(pending)
        if len(nodes) >= N or len(nodes) == prev:
            break
        if len(nodes) + 3 * n_expand > N:
            stop = False
            while not stop:
                prev = len(nodes)
                order = sorted(pending, key=lambda m: (len(m["keys"]), m["serial"]))
                pending = []
                for n in reversed(order):                  # largest first; among equals the one created last
                    expand(n, nodes, pending)
                    if len(nodes) >= N:
                        break
                stop = len(nodes) >= N or len(nodes) == prev
            break
    out = []
    for n in nodes:
        best = n["keys"][0]
        for k in n["keys"][1:]:
            if cand[k][2] > cand[best][2]:
                best = k
        out.append(best)
    return out


@pytest.mark.parametrize("seed,shape,nfeat", [(5, (192, 640), 1000), (6, (240, 320), 2500), (7, (375, 1242), 2500)])
def test_python_quadtree_matches_oracle_extraction(seed, shape, nfeat):
    rng = np.random.default_rng(seed)
    H, W = shape
    base = rng.normal(size=(H // 4 + 2, W // 4 + 2))
    img = np.kron(base, np.ones((4, 4)))[:H, :W] * 25 + rng.normal(size=(H, W)) * 12 + 128
    for _ in range(600):                                    # bright blobs: corners with distinct responses
        y, x = rng.integers(8, H - 8), rng.integers(8, W - 8)
        img[y - 1:y + 2, x - 1:x + 2] += rng.uniform(40, 110)
    img = np.clip(img, 0, 255).astype(np.uint8)
    p = ol.default_orb_params(nfeat)
    kps = ol.orb_extract(img, p)
    pyr = ol.orb_pyramid(img, p)
    quota = ol.level_quotas(p)
    _, _, scale = ol.level_sizes(W, H, p)
    checked = 0
    for l, lvl in enumerate(pyr):
        h, w = lvl.shape
        minB, maxBX, maxBY = EDGE - 3, w - EDGE + 3, h - EDGE + 3
        xs, ys, sc = ol.level_candidates(lvl, p)
        cand = [(float(x), float(y), float(s)) for x, y, s in zip(xs, ys, sc)]
        keep = distribute(cand, minB, maxBX, minB, maxBY, int(quota[l]))
        want = kps[kps["octave"] == l]
        assert len(keep) == len(want), f"level {l}: {len(keep)} vs {len(want)}"
        for k, ref in zip(keep, want):
            x = np.float32(cand[k][0] + minB); y = np.float32(cand[k][1] + minB)
            if l > 0:
                x = np.float32(x * scale[l]); y = np.float32(y * scale[l])
            assert (x, y, np.float32(cand[k][2])) == (ref["x"], ref["y"], ref["response"]), f"level {l}"
            checked += 1
    assert checked == len(kps) and checked > 200
