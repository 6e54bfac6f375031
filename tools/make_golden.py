"""Generates tests/golden/*.npz with the CPU oracle (oracle/pt_oracle.c).

The reference ships no golden vectors (parity unpinned), so these are
produced by the brute-force restatement of src/Distance.h:6-11 and cross-checked here against
scipy's cKDTree before being written.  Re-run: `python tools/make_golden.py`.
"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pto  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def pack(xyz, rng):
    n = xyz.shape[0]
    nrm = rng.standard_normal((n, 3)).astype(np.float32)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    col = rng.integers(0, 256, (n, 3)).astype(np.int32)
    return nrm.astype(np.float32), col


def case(name, xyz, q, ks, radius=None, rng=None, check_scipy=True):
    rng = rng or np.random.default_rng(1)
    nrm, col = pack(xyz, rng)
    P = pto.make_points(xyz, normal=nrm, color=col)
    Q = pto.make_points(q)
    out = {"xyz": xyz, "normal": nrm, "color": col, "queries": q,
           "radius": np.float64(-1.0 if radius is None else radius), "ks": np.array(ks)}
    for k in ks:
        idx, d2 = pto.knn_bruteforce(P, Q, k, radius=-1.0 if radius is None else radius)
        rgba, nout = pto.blend(P, idx, d2)
        if check_scipy and radius is None and xyz.shape[0] >= k:
            sd, si = cKDTree(xyz).query(q, k)
            sd = sd.reshape(len(q), k)
            # distances must agree everywhere; indices wherever the distance is unique
            assert np.allclose(np.sqrt(d2), sd, rtol=1e-12, atol=0), name
        out[f"idx_k{k}"] = idx
        out[f"d2_k{k}"] = d2
        out[f"rgba_k{k}"] = rgba
        out[f"normal_k{k}"] = nout
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, xyz.shape, q.shape, ks)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    # 1. random cloud, fp32-representable
    xyz = f32(rng.random((4096, 3)) * 10.0)
    q = f32(rng.random((256, 3)) * 10.0)
    case("random_f32", xyz, q, [1, 8, 20, 32], rng=rng)
    # 2. same cloud, coordinates NOT fp32-representable (forces the fp64 records)
    xyz64 = rng.random((4096, 3)) * 10.0
    q64 = rng.random((256, 3)) * 10.0
    case("random_f64", xyz64, q64, [1, 8, 20, 32], rng=rng)
    # 3. integer lattice: massive exact distance ties -> lowest index must win
    g = np.stack(np.meshgrid(*[np.arange(10.0)] * 3, indexing="ij"), -1).reshape(-1, 3)
    g = g[rng.permutation(len(g))]
    case("lattice_ties", g, g[::5].copy(), [1, 8, 20, 32], rng=rng, check_scipy=False)
    # 4. duplicated points (position duplicates; Point::operator== compares position only)
    base = f32(rng.random((300, 3)))
    dup = np.concatenate([base, base[:150], base[:50]])
    dup = dup[rng.permutation(len(dup))]
    case("duplicates", dup, f32(rng.random((64, 3))), [1, 8, 20], rng=rng, check_scipy=False)
    # 5. k > N: short lists padded with -1 / +inf
    tiny = f32(rng.random((5, 3)))
    case("k_gt_n", tiny, f32(rng.random((16, 3))), [8, 32], rng=rng, check_scipy=False)
    # 6. collinear / coplanar degenerate clouds
    t = f32(rng.random(2000))
    line = np.stack([t, f32(2 * t), f32(-t)], 1)
    case("collinear", line, f32(rng.random((64, 3))), [1, 8, 20], rng=rng, check_scipy=False)
    plane = np.stack([f32(rng.random(3000)), f32(rng.random(3000)), np.zeros(3000)], 1)
    case("coplanar", plane, f32(rng.random((64, 3))), [8, 20], rng=rng)
    # 7. radius-bounded with empty and short results
    xyz = f32(rng.random((4096, 3)) * 10.0)
    q = f32(rng.random((256, 3)) * 12.0 - 1.0)
    case("radius_bounded", xyz, q, [8, 16], radius=0.45, rng=rng, check_scipy=False)
    # 8. queries that coincide with cloud points (d2 == 0 exact hits)
    xyz = f32(rng.random((2048, 3)))
    case("exact_hits", xyz, xyz[::16].copy(), [1, 8, 20], rng=rng)
    # 9. single point / clustered extremes
    clus = np.concatenate([f32(rng.normal(0, 1e-3, (2000, 3))), f32(rng.random((100, 3)) * 100)])
    case("skewed", clus, f32(rng.random((64, 3)) * 100), [8, 16, 32], rng=rng)


if __name__ == "__main__":
    main()
