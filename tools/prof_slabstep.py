"""Prices the pieces of the multi-GPU steady-state step (ghost-zone path) on ONE GPU:
the owner query with ids + d2, the ghost check, and the small torch ops around them."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
w = pkg.synth.CONFIGS["cfg2"]
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv); m, k = q.shape[0], w.k
ids = torch.arange(w.n_points, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tree = pkg.DeviceTree(pos, attrs, ids)
eng = pkg.dist.CudaSlabEngine(tree)
slab = pkg.dist.SlabTransfer(eng, own_box=pkg.dist.points_box(pos), halo=2.0)


def timeit(name, fn, n=13):
    ts = []
    for it in range(n):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if it >= 3: ts.append(a.elapsed_time(b))
    print(f"{name}: {sum(ts)/len(ts):.4f} ms  (min {min(ts):.4f})")


idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
nrm = torch.empty((m, 3), dtype=torch.float32, device=dev); d2 = torch.empty((m, k), dtype=torch.float64, device=dev)
timeit("query ids+d2 into fixed buffers", lambda: tree.query(q, k, idx=idx, d2=d2, rgba=rgba, normal=nrm))
timeit("engine.query (fresh outputs)", lambda: eng.query(q, k, outputs=True, want_d2=True, want_cand=False))
timeit("_transfer_ghost validate=False", lambda: slab._transfer_ghost(q, k, None, False, validate=False))
timeit("transfer validate=False", lambda: slab.transfer(q, k, validate=False))
viol = torch.zeros((1,), dtype=torch.int32, device=dev)
timeit("ghost_check only", lambda: eng.ghost_check(q, d2, k, None, slab.boxes6, 0, 2.0, viol))
tree.close()
