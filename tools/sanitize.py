"""Small end-to-end pass over every kernel family (index build, every query kernel, the slab /
halo kernels, the texture stage), written to be run under `compute-sanitizer --tool memcheck`.
compute-sanitizer is closed on this round's GPU pool, so the pass is run plain, with the pool's
guard words switched on instead (option "pool_guard", see pt_build.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge

pkg = ge.package()
torch.cuda.set_device(0)
pkg.set_option("pool_guard", 1)
side = 20.0
P = pkg.synth.cloud_host(30_000, seed=3, side=side)
V = pkg.synth.samples_host(24, side=side)
V["U"] = V["ver"][:, 0] / side * 0.9 + 0.05
V["V"] = V["ver"][:, 1] / side * 0.9 + 0.05
F = pkg.synth.grid_faces(24, 24)
for order in (1, 2):
    pkg.set_option("order", order)
    for variant in (-1, 6, 5, 2, 0):
        pkg.set_option("knn_variant", variant)
        with pkg.Tree(P) as t:
            out = t.transfer(V, 20, want_idx=True, want_d2=True)
            t.knn(V, 8, radius=0.5)
pkg.set_option("order", 1)
pkg.set_option("knn_variant", -1)
with pkg.Tree(P) as t:
    img, st = t.texture(V, F, k=20, resolution=256)
with pkg.ShardedTree(P, [0, 0, 0]) as s:
    o2 = s.transfer(V, 20, want_idx=True)
    assert np.array_equal(o2["idx"], out["idx"])
print("sanitize pass done:", st)
print("pool_guard_hits", pkg.get_option("pool_guard_hits"))
assert pkg.get_option("pool_guard_hits") == 0
