"""CPU simulation of the leaf/box pyramid to size design parameters before spending GPU time.
Counts, for exact kNN with the ideal bound (true k-th d2), how many 32-point leaves and
how many 32-wide box groups a query must touch, for Morton vs Hilbert point order."""
import sys, time
import numpy as np
from scipy.spatial import cKDTree

def surface(u, v):
    z = np.zeros_like(u)
    for a, f, g, p, q in [(12.0, 0.021, 0.017, 0.3, 1.1), (6.0, 0.047, 0.039, 1.7, 0.2),
                          (2.5, 0.11, 0.13, 2.9, 4.1), (0.8, 0.31, 0.27, 0.5, 3.3)]:
        z += a * np.sin(f * u + p) * np.sin(g * v + q)
    return z

def part1by2(x):
    x = x.astype(np.uint64) & np.uint64(0x1fffff)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1f00000000ffff)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x

def morton(c):
    return part1by2(c[:, 0]) | (part1by2(c[:, 1]) << np.uint64(1)) | (part1by2(c[:, 2]) << np.uint64(2))

def hilbert3(c, bits=21):
    # Skilling's transpose algorithm (AxestoTranspose), vectorised
    X = [c[:, i].astype(np.uint64).copy() for i in range(3)]
    M = np.uint64(1) << np.uint64(bits - 1)
    Q = M
    while Q > np.uint64(1):
        P = Q - np.uint64(1)
        for i in range(3):
            sel = (X[i] & Q) != 0
            X[0] = np.where(sel, X[0] ^ P, X[0])
            t = (X[0] ^ X[i]) & P
            t = np.where(sel, np.uint64(0), t)
            X[0] ^= t
            X[i] ^= t
        Q >>= np.uint64(1)
    for i in range(1, 3):
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > np.uint64(1):
        t = np.where((X[2] & Q) != 0, t ^ (Q - np.uint64(1)), t)
        Q >>= np.uint64(1)
    for i in range(3):
        X[i] ^= t
    # interleave: X[0] is most significant
    return (part1by2(X[0]) << np.uint64(2)) | (part1by2(X[1]) << np.uint64(1)) | part1by2(X[2])

def lb(q, lo, hi):
    e = np.maximum(np.maximum(lo - q, q - hi), 0.0)
    return (e * e).sum(-1)

def run(order, n=2_000_000, side=200.0, k=16, spacing=2.24, leaf=32, W=32, seed=0):
    rng = np.random.default_rng(seed)
    u = rng.random(n) * side; v = rng.random(n) * side
    z = surface(u, v) + rng.normal(0, 0.01, n)
    P = np.stack([u, v, z], 1).astype(np.float32).astype(np.float64)
    lo = P.min(0); ext = (P.max(0) - lo).max()
    # emulate full-size resolution: cell = 1000/2^21
    cell = 1000.0 / 2**21
    c = np.minimum(((P - lo) / cell).astype(np.int64), 2**21 - 1)
    key = morton(c) if order == "morton" else hilbert3(c)
    perm = np.argsort(key, kind="stable")
    P = P[perm]
    nl = (n + leaf - 1) // leaf
    pad = nl * leaf - n
    Pp = np.concatenate([P, np.repeat(P[-1:], pad, 0)]) if pad else P
    L = Pp.reshape(nl, leaf, 3)
    levels = [(L.min(1), L.max(1))]
    while levels[-1][0].shape[0] > 1:
        lo_, hi_ = levels[-1]
        cnt = lo_.shape[0]; g = (cnt + W - 1) // W
        padn = g * W - cnt
        if padn:
            lo_ = np.concatenate([lo_, np.full((padn, 3), np.inf)]); hi_ = np.concatenate([hi_, np.full((padn, 3), -np.inf)])
        levels.append((lo_.reshape(g, W, 3).min(1), hi_.reshape(g, W, 3).max(1)))
    gx = np.arange(5.0, side - 5.0, spacing)
    qu, qv = np.meshgrid(gx, gx); qu = qu.ravel(); qv = qv.ravel()
    Q = np.stack([qu, qv, surface(qu, qv)], 1).astype(np.float32).astype(np.float64)
    Q = Q[:: max(1, len(Q) // 1500)]
    d, _ = cKDTree(P).query(Q, k)
    kth = d[:, -1] ** 2
    leaves = []; groups = [[] for _ in levels]
    for q, b in zip(Q, kth):
        # top-down with ideal bound
        nodes = np.arange(levels[-1][0].shape[0])
        for li in range(len(levels) - 1, -1, -1):
            lo_, hi_ = levels[li]
            # children group loads at this level = number of parent nodes expanded
            ok = lb(q, lo_[nodes], hi_[nodes]) <= b
            nodes = nodes[ok]
            if li > 0:
                groups[li - 1].append(len(nodes))  # each surviving node => one group load at level li-1
                cnt = levels[li - 1][0].shape[0]
                nodes = (nodes[:, None] * W + np.arange(W)[None, :]).ravel()
                nodes = nodes[nodes < cnt]
        leaves.append(len(nodes))
    leaves = np.array(leaves)
    print(f"{order:8s} leaf={leaf} W={W} k={k}: leaves/query mean {leaves.mean():.2f} p95 {np.percentile(leaves,95):.0f} max {leaves.max()}"
          f" | cand/query {leaves.mean()*leaf:.0f} | group loads per level (bottom first): "
          + " ".join(f"{np.mean(g):.2f}" for g in groups if g))

if __name__ == "__main__":
    for order in ("morton", "hilbert"):
        for leaf, W in ((32, 32), (16, 16), (8, 8), (32, 8), (32, 4)):
            run(order, leaf=leaf, W=W)
