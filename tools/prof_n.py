"""Query time of the cfg2 samples against prefixes of the cfg2 cloud (address/size sensitivity)."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
w = pkg.synth.CONFIGS["cfg2"]; n = w.n_points; k = w.k
pos, attrs = pkg.synth.cloud_device(n, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv); m = q.shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
nrm = torch.empty((m, 3), dtype=torch.float32, device=dev)
import os
if os.environ.get("PT_VARIANT"): pkg.set_option("knn_variant", int(os.environ["PT_VARIANT"]))
out = []
for nn in [int(a) for a in sys.argv[1:]] or [n, 49900000, 49000000, 50000000 - 32 * 7]:
    t = pkg.DeviceTree(pos[:nn], attrs[:nn])
    ts = []
    for it in range(13):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); t.query(q, k, idx=idx, rgba=rgba, normal=nrm); e1.record(); torch.cuda.synchronize()
        if it >= 3: ts.append(e0.elapsed_time(e1))
    out.append(f"n={nn}: {sum(ts)/len(ts):.4f}")
    t.close()
print("  ".join(out), flush=True)
