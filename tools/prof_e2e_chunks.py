"""End-to-end pt_transfer (pinned host buffers) against the number of pipeline chunks.
usage: prof_e2e_chunks.py [chunks ...]"""
import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0)
w = pkg.synth.CONFIGS["cfg2"]
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed, kind=w.kind, sigma=w.sigma)
q = pkg.synth.samples_device(w.gu, w.gv, center=w.center)
m, k = q.shape[0], w.k
tree = pkg.DeviceTree(pos, attrs)
qh = pkg.synth.queries_to_host(q, pinned=True).numpy().view(pkg.POINT_DTYPE).reshape(-1)
out = {"idx": torch.empty((m, k), dtype=torch.int32, pin_memory=True).numpy(),
       "rgba": torch.empty((m, 4), dtype=torch.uint8, pin_memory=True).numpy(),
       "normal": torch.empty((m, 3), dtype=torch.float32, pin_memory=True).numpy()}
for chunks in [int(x) for x in sys.argv[1:]] or [2, 4, 6, 8, 12, 16, 24]:
    pkg.set_option("host_chunks", chunks)
    for _ in range(3):
        tree.transfer(qh, k, out=out)
    t0 = time.perf_counter()
    for _ in range(20):
        tree.transfer(qh, k, out=out)
    print(f"chunks {chunks:3d}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms", flush=True)
