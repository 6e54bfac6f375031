"""Two kernel variants against each other on a workload (default 2 vs 5 on cfg2): bit-equality of every output, timing,
and (with a -DPT_STATS build) the work counters."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = pkg.synth.CONFIGS[cfg]; k = w.k
n = min(w.n_points, 60_000_000)
pos, attrs = pkg.synth.cloud_device(n, w.seed, kind=w.kind, sigma=w.sigma)
q = pkg.synth.samples_device(w.gu, w.gv, center=w.center); m = q.shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tree = pkg.DeviceTree(pos, attrs)
res = {}
variants = [int(a) for a in sys.argv[2:]] or [2, 5]
for variant in variants:
    pkg.set_option("knn_variant", variant)
    idx = torch.full((m, k), -7, dtype=torch.int32, device=dev); rgba = torch.zeros((m, 4), dtype=torch.uint8, device=dev)
    nrm = torch.zeros((m, 3), dtype=torch.float32, device=dev); d2 = torch.zeros((m, k), dtype=torch.float64, device=dev)
    tree.query(q, k, radius=w.radius, idx=idx, d2=d2, rgba=rgba, normal=nrm); torch.cuda.synchronize()
    res[variant] = (idx.clone(), d2.clone(), rgba.clone(), nrm.clone())
    pkg.api.debug_stats()
    ts = []
    for it in range(13):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); tree.query(q, k, radius=w.radius, idx=idx, rgba=rgba, normal=nrm); e1.record(); torch.cuda.synchronize()
        if it >= 3: ts.append(e0.elapsed_time(e1))
    st = pkg.api.debug_stats()
    per = {kk: round(v / 13 / m, 3) for kk, v in st.items()} if st["samples"] else {}
    print(f"{cfg} variant {variant}: {sum(ts)/len(ts):.4f} ms (min {min(ts):.4f})  m={m} k={k} "
          f"fallback={tree.info().last_fallback_samples} {per}", flush=True)
a, b = res[variants[0]], res[variants[-1]]
for name, x, y in zip(("idx", "d2", "rgba", "normal"), a, b):
    same = bool(torch.equal(x, y))
    print(f"  {name}: {'identical' if same else 'DIFFERENT: %d rows' % int((x != y).reshape(m, -1).any(dim=1).sum())}")
tree.close()
