"""SAH cost of the LBVH the library builds (via the oracle's bit-exact host rebuild) against a full-sweep SAH build,
SAH splits over the Morton order, and tree rotations — the numbers DESIGN.md §4 / §7 quote.  CPU only.

    python tools/bvh_quality.py            # C4 cover scene and the C5 spot mesh (+ Cornell quads)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from rendering_learning_b200 import ow, scenes  # noqa: E402

sys.setrecursionlimit(20000)


def area(lo, hi):
    d = np.maximum(hi - lo, 0)
    return 2 * (d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 0] * d[..., 2])


def world_aabbs(world, skip_big_spheres=True):
    """world-space boxes of the leaves, transforms applied like csrc/flatten.cpp does"""
    boxes = []

    def walk(h, M, t):
        if isinstance(h, ow.Sphere):
            if skip_big_spheres and h.radius >= 64:
                return
            c1 = np.array(h.center[1])
            c2 = np.array(h.center[2]) if h.center[0] == "moving" else c1
            c1, c2 = M @ c1 + t, M @ c2 + t
            boxes.append(np.concatenate([np.minimum(c1, c2) - h.radius, np.maximum(c1, c2) + h.radius]))
        elif isinstance(h, ow.Triangle):
            P = np.array(h.points) @ M.T + t
            boxes.append(np.concatenate([P.min(0), P.max(0)]))
        elif isinstance(h, ow.Quad):
            q, u, v = np.array(h.q), np.array(h.u), np.array(h.v)
            P = np.array([q, q + u, q + v, q + u + v]) @ M.T + t
            boxes.append(np.concatenate([P.min(0) - 1e-4, P.max(0) + 1e-4]))
        elif isinstance(h, ow.Transform):
            walk(h.object, M @ np.array(h.m), t)
        elif isinstance(h, ow.Translate):
            walk(h.object, M, t + M @ np.array(h.offset))
        elif hasattr(h, "children"):
            for c in h.children:
                walk(c, M, t)
        elif isinstance(h, (list, tuple)):
            for c in h:
                walk(c, M, t)

    walk(world, np.eye(3), np.zeros(3))
    return np.array(boxes, np.float32)


def lbvh_cost(aabb):
    h = orc.lbvh_build(aabb)
    nb = h["node_aabb"]
    return float(area(nb[:, :3], nb[:, 3:]).sum() / area(nb[0, :3], nb[0, 3:])), h


def sweep_sah_cost(aabb, order=None):
    """top-down build; order=None: full sweep over the three centroid axes, else SAH splits over the given fixed order"""
    cent = (aabb[:, :3] + aabb[:, 3:]) * 0.5
    cost, root = [0.0], [None]

    def rec(idx):
        lo, hi = aabb[idx, :3].min(0), aabb[idx, 3:].max(0)
        s = area(lo, hi)
        if root[0] is None:
            root[0] = s
        if len(idx) == 1:
            return
        cost[0] += s / root[0]
        best = (np.inf, None, None)
        orders = [idx] if order is not None else [idx[np.argsort(cent[idx, ax], kind="stable")] for ax in range(3)]
        for o in orders:
            n = len(o)
            lo_l, hi_l = np.minimum.accumulate(aabb[o, :3], 0), np.maximum.accumulate(aabb[o, 3:], 0)
            lo_r = np.minimum.accumulate(aabb[o[::-1], :3], 0)[::-1]
            hi_r = np.maximum.accumulate(aabb[o[::-1], 3:], 0)[::-1]
            k = np.arange(1, n)
            c = area(lo_l[:-1], hi_l[:-1]) * k + area(lo_r[1:], hi_r[1:]) * (n - k)
            j = int(np.argmin(c))
            if c[j] < best[0]:
                best = (c[j], o, j + 1)
        _, o, m = best
        rec(o[:m])
        rec(o[m:])

    rec(np.arange(len(aabb)) if order is None else np.asarray(order))
    return cost[0]


def rotation_cost(aabb, passes=4):
    """Kensler-style rotations (swap a child with a grandchild of its sibling) applied bottom-up on the LBVH"""
    _, h = lbvh_cost(aabb)
    n = len(aabb)
    left, right, sp = h["left"].copy(), h["right"].copy(), h["sorted_prim"]
    nb = h["node_aabb"].astype(np.float64).copy()
    box = lambda c: nb[c] if c >= 0 else aabb[sp[~c]].astype(np.float64)
    merge = lambda a, b: np.concatenate([np.minimum(a[:3], b[:3]), np.maximum(a[3:], b[3:])])
    sa = lambda b: float(area(b[:3], b[3:]))
    total = lambda: sum(sa(nb[i]) for i in range(n - 1)) / sa(nb[0])

    def post_order():
        out, st = [], [(0, False)]
        while st:
            x, done = st.pop()
            if x < 0:
                continue
            if done:
                out.append(x)
                continue
            st += [(x, True), (left[x], False), (right[x], False)]
        return out

    costs = [total()]
    for _ in range(passes):
        for x in post_order():
            best, move = 0.0, None
            for a, b, side in ((left[x], right[x], 0), (right[x], left[x], 1)):
                if b < 0:
                    continue
                for g, o, which in ((left[b], right[b], 0), (right[b], left[b], 1)):
                    gain = sa(nb[b]) - sa(merge(box(a), box(o)))
                    if gain > best + 1e-12:
                        best, move = gain, (a, b, side, g, o)
            if move:
                a, b, side, g, o = move
                left[b], right[b] = a, o
                nb[b] = merge(box(a), box(o))
                if side == 0:
                    left[x] = g
                else:
                    right[x] = g
        costs.append(total())
    return costs


if __name__ == "__main__":
    orc.build()
    cover = world_aabbs(scenes.ow_cover_world())
    cow = world_aabbs(scenes.ow_cow_world())
    for name, a in (("C4 cover scene (487 spheres, ground on the big list)", cover), ("C5 Cornell quads + spot (5862)", cow),
                    ("C5 spot mesh alone (5856, quads on the big list)", cow[6:])):
        c, h = lbvh_cost(a)
        print(f"{name}\n  LBVH (cubic Morton cells)      SAH cost {c:8.3f}")
        print(f"  SAH splits over Morton order   SAH cost {sweep_sah_cost(a, h['sorted_prim']):8.3f}")
        print(f"  full-sweep SAH build           SAH cost {sweep_sah_cost(a):8.3f}")
        print("  LBVH + tree rotations          SAH cost " + " -> ".join(f"{x:.3f}" for x in rotation_cost(a)))
