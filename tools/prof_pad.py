"""Query time of one kernel variant on cfg2 over a range of smem pads (blocks per SM)."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
variant = int(sys.argv[1]); pads = [int(a) for a in sys.argv[2:]] or [0]
w = pkg.synth.CONFIGS["cfg2"]; k = w.k
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv); m = q.shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pkg.set_option("knn_variant", variant)
tree = pkg.DeviceTree(pos, attrs)
idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
nrm = torch.empty((m, 3), dtype=torch.float32, device=dev)
out = []
for pad in pads:
    pkg.set_option("smem_pad", pad)
    pkg.api.debug_stats()
    ts = []
    for it in range(13):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); tree.query(q, k, idx=idx, rgba=rgba, normal=nrm); e1.record(); torch.cuda.synchronize()
        if it >= 3: ts.append(e0.elapsed_time(e1))
    st = pkg.api.debug_stats()
    out.append(f"pad={pad}: {sum(ts)/len(ts):.4f}" + (f" ovf/launch={st['overflowed']/13:.1f}" if st["samples"] else ""))
print(f"variant {variant}: " + "  ".join(out), flush=True)
pkg.set_option("smem_pad", 0); tree.close()
