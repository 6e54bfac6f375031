"""CPU simulation of traversal policies (dynamic bound) on the leaf/box pyramid:
counts leaves scanned / node groups expanded / heap insertions per query."""
import sys, heapq
import numpy as np
sys.path.insert(0, "/root/repo/tools")
from sim_tree import surface, morton, hilbert3, lb
from scipy.spatial import cKDTree

def build(order, n, side, leaf, B, seed=0):
    rng = np.random.default_rng(seed)
    u = rng.random(n) * side; v = rng.random(n) * side
    z = surface(u, v) + rng.normal(0, 0.01, n)
    P = np.stack([u, v, z], 1).astype(np.float32).astype(np.float64)
    lo = P.min(0)
    cell = 1000.0 / 2**21
    c = np.minimum(((P - lo) / cell).astype(np.int64), 2**21 - 1)
    if order == "morton": key = morton(c)
    elif order == "hilbert": key = hilbert3(c)
    elif order == "morton2d": 
        from sim_tree import part1by2
        key = part1by2(c[:,0]) | (part1by2(c[:,1]) << np.uint64(1))
    perm = np.argsort(key, kind="stable"); P = P[perm]
    nl = n // leaf; P = P[: nl * leaf]
    L = P.reshape(nl, leaf, 3)
    levels = [(L.min(1), L.max(1))]
    while levels[-1][0].shape[0] > 1:
        lo_, hi_ = levels[-1]; cnt = lo_.shape[0]; g = (cnt + B - 1) // B; pad = g * B - cnt
        if pad:
            lo_ = np.concatenate([lo_, np.full((pad, 3), np.inf)]); hi_ = np.concatenate([hi_, np.full((pad, 3), -np.inf)])
        levels.append((lo_.reshape(g, B, 3).min(1), hi_.reshape(g, B, 3).max(1)))
    return P, L, levels

def query_dfs(q, k, L, levels, B, policy):
    """near-first DFS with dynamic bound. policy: 'lb' | 'center' (tie/ordering heuristic)"""
    top = len(levels) - 1
    best = []  # max-heap of (-d2)
    bound = np.inf
    stats = dict(leaves=0, expands=0, inserts=0, boxtests=0)
    def order_children(li, ids):
        lo_, hi_ = levels[li]
        l = lb(q, lo_[ids], hi_[ids]); stats['boxtests'] += len(ids)
        if policy == 'lb': keyv = l
        else:
            c = 0.5 * (lo_[ids] + hi_[ids]); keyv = l + 1e-3 * ((c - q) ** 2).sum(-1)
        o = np.argsort(keyv, kind="stable")
        return ids[o], l[o]
    def rec(li, ids):
        nonlocal bound
        ids, ls = order_children(li, ids)
        stats['expands'] += 1
        for i, l in zip(ids, ls):
            if l > bound: 
                if policy == 'lb': break
                continue
            if li == 0:
                stats['leaves'] += 1
                d = ((L[i] - q) ** 2).sum(-1)
                for dd in d:
                    if len(best) < k: heapq.heappush(best, -dd); stats['inserts'] += 1
                    elif dd < -best[0]: heapq.heapreplace(best, -dd); stats['inserts'] += 1
                if len(best) == k: bound = -best[0]
            else:
                cnt = levels[li - 1][0].shape[0]
                ch = np.arange(i * B, min(i * B + B, cnt))
                rec(li - 1, ch)
    rec(top, np.arange(levels[top][0].shape[0]))
    return stats

def query_bestfirst(q, k, L, levels, B):
    top = len(levels) - 1
    pq = []
    stats = dict(leaves=0, expands=0, inserts=0, boxtests=0, pushes=0)
    best = []; bound = np.inf
    def push(li, ids):
        lo_, hi_ = levels[li]; l = lb(q, lo_[ids], hi_[ids]); stats['boxtests'] += len(ids)
        for i, ll in zip(ids, l):
            if ll <= bound: heapq.heappush(pq, (ll, li, int(i))); stats['pushes'] += 1
    push(top, np.arange(levels[top][0].shape[0]))
    while pq:
        l, li, i = heapq.heappop(pq)
        if l > bound: break
        if li == 0:
            stats['leaves'] += 1
            d = ((L[i] - q) ** 2).sum(-1)
            for dd in d:
                if len(best) < k: heapq.heappush(best, -dd); stats['inserts'] += 1
                elif dd < -best[0]: heapq.heapreplace(best, -dd); stats['inserts'] += 1
            if len(best) == k: bound = -best[0]
        else:
            stats['expands'] += 1
            cnt = levels[li - 1][0].shape[0]
            push(li - 1, np.arange(i * B, min(i * B + B, cnt)))
    return stats

def run(order, leaf, B, k=16, n=1_000_000, side=141.4, spacing=2.24, nq=300):
    P, L, levels = build(order, n, side, leaf, B)
    gx = np.arange(5.0, side - 5.0, spacing); qu, qv = np.meshgrid(gx, gx)
    Q = np.stack([qu.ravel(), qv.ravel(), surface(qu.ravel(), qv.ravel())], 1).astype(np.float32).astype(np.float64)
    Q = Q[:: max(1, len(Q) // nq)]
    for name, fn in (("dfs-lb", lambda q: query_dfs(q, k, L, levels, B, 'lb')),
                     ("dfs-center", lambda q: query_dfs(q, k, L, levels, B, 'center')),
                     ("best-first", lambda q: query_bestfirst(q, k, L, levels, B))):
        S = [fn(q) for q in Q]
        keys = S[0].keys()
        print(f"{order:8s} leaf={leaf:2d} B={B:2d} {name:11s} " + " ".join(f"{kk}={np.mean([s[kk] for s in S]):.1f}" for kk in keys))

if __name__ == "__main__":
    for order in ("morton", "hilbert"):
        for leaf, B in ((32, 32), (32, 8)):
            run(order, leaf, B)
