"""Cost of the optional outputs / id mapping of the query kernel (1 GPU)."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
w = pkg.synth.CONFIGS["cfg2"]
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv); m, k = q.shape[0], w.k
ids = torch.arange(w.n_points, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for use_ids in (False, True):
    tree = pkg.DeviceTree(pos, attrs, ids if use_ids else None)
    for want_d2, want_cand in ((False, False), (True, False), (False, True)):
        idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
        nrm = torch.empty((m, 3), dtype=torch.float32, device=dev)
        d2 = torch.empty((m, k), dtype=torch.float64, device=dev) if want_d2 else None
        cand = torch.empty((m * k * 32,), dtype=torch.uint8, device=dev) if want_cand else None
        ts = []
        for it in range(13):
            flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); tree.query(q, k, idx=idx, d2=d2, rgba=rgba, normal=nrm, cand=cand); b.record(); torch.cuda.synchronize()
            if it >= 3: ts.append(a.elapsed_time(b))
        print(f"ids={use_ids} d2={want_d2} cand={want_cand}: {sum(ts)/len(ts):.4f} ms")
    tree.close()
