"""Throughput of kernel variants on the cfg2 cloud at several sample counts (wave effects)."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
variants = [int(a) for a in sys.argv[1:]] or [2, 5]
import os
GRIDS = tuple(int(g) for g in os.environ.get("PT_GRIDS", "448,896").split(","))
w = pkg.synth.CONFIGS["cfg2"]; k = w.k
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tree = pkg.DeviceTree(pos, attrs)
for g in GRIDS:
    q = pkg.synth.samples_device(g, g); m = q.shape[0]
    idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
    nrm = torch.empty((m, 3), dtype=torch.float32, device=dev)
    line = f"m={m:8d} blocks={-(-m//32):6d}"
    for v in variants:
        pkg.set_option("knn_variant", v)
        ts = []
        for it in range(9):
            flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); tree.query(q, k, idx=idx, rgba=rgba, normal=nrm); e1.record(); torch.cuda.synchronize()
            if it >= 3: ts.append(e0.elapsed_time(e1))
        t = sum(ts) / len(ts)
        line += f"  v{v}: {t:.4f} ms {m / t / 1e3:7.1f} Msamples/s fb={tree.info().last_fallback_samples}"
    print(line, flush=True)
tree.close()
