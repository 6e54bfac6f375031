"""End-to-end time of pt_transfer (pinned host buffers in and out) on cfg2 vs the number of
pipeline chunks of the host-buffer API."""
import sys, time; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
w = pkg.synth.CONFIGS["cfg2"]; k = w.k
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv); m = q.shape[0]
import os
if os.environ.get("PT_VARIANT"): pkg.set_option("knn_variant", int(os.environ["PT_VARIANT"]))
tree = pkg.DeviceTree(pos, attrs)
q_host = pkg.synth.queries_to_host(q, pinned=True)
qh = q_host.numpy().view(pkg.POINT_DTYPE).reshape(-1)
out = {"idx": torch.empty((m, k), dtype=torch.int32, pin_memory=True).numpy(),
       "rgba": torch.empty((m, 4), dtype=torch.uint8, pin_memory=True).numpy(),
       "normal": torch.empty((m, 3), dtype=torch.float32, pin_memory=True).numpy()}
for chunks in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8, 12, 16]:
    pkg.set_option("host_chunks", chunks)
    for _ in range(3): tree.transfer(qh, k, out=out)
    t0 = time.perf_counter()
    for _ in range(20): tree.transfer(qh, k, out=out)
    t = (time.perf_counter() - t0) / 20
    print(f"chunks={chunks:2d}: {t*1e3:.4f} ms  ({m/t/1e6:.1f} Msamples/s)  device {tree.info().last_query_ms:.4f} ms", flush=True)
tree.close()
