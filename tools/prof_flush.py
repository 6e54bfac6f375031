"""Does the L2-flush method interact with the measured kernel time?  cfg2 query against two
cloud sizes under (a) a 256 MiB write before every pass (bench.py's method), (b) no flush,
(c) a 256 MiB read (sum) before every pass."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
w = pkg.synth.CONFIGS["cfg2"]; n = w.n_points; k = w.k
pos, attrs = pkg.synth.cloud_device(n, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv); m = q.shape[0]
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
sink = torch.zeros(1, dtype=torch.int64, device=dev)
idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
nrm = torch.empty((m, 3), dtype=torch.float32, device=dev)
for nn in (n, 49999776, 49900000):
    t = pkg.DeviceTree(pos[:nn], attrs[:nn])
    line = f"n={nn}:"
    for mode in ("write", "none", "read"):
        ts = []
        for it in range(13):
            if mode == "write": flush.zero_()
            elif mode == "read": sink.copy_(flush.view(torch.int64).sum().view(1))
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); t.query(q, k, idx=idx, rgba=rgba, normal=nrm); e1.record(); torch.cuda.synchronize()
            if it >= 3: ts.append(e0.elapsed_time(e1))
        line += f"  {mode} {sum(ts)/len(ts):.4f}"
    print(line, flush=True)
    t.close()
