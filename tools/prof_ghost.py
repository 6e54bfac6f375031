"""Why is the ghost-augmented index slower?  One GPU: the cfg2 cloud plus ~95k extra points
placed (a) in a strip just outside the domain (what exchange_ghosts adds), (b) inside it.
With a -DPT_STATS build the per-sample work counters are printed as well."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
w = pkg.synth.CONFIGS["cfg2"]; L = pkg.synth.L_DOMAIN; n = w.n_points; k = w.k
pos, attrs = pkg.synth.cloud_device(n, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv); m = q.shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
nrm = torch.empty((m, 3), dtype=torch.float32, device=dev)


def run(name, p, a, qq=q):
    t = pkg.DeviceTree(p, a)
    mm = qq.shape[0]
    ts = []
    pkg.api.debug_stats()
    for it in range(13):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); t.query(qq, k, idx=idx[:mm], rgba=rgba[:mm], normal=nrm[:mm]); e1.record(); torch.cuda.synchronize()
        if it >= 3: ts.append(e0.elapsed_time(e1))
    st = pkg.api.debug_stats()
    per = {kk: round(v / 13 / mm, 3) for kk, v in st.items()} if st["samples"] else {}
    if per: per["overflowed_per_launch"] = st["overflowed"] / 13; per["compactions_per_launch"] = st["compactions"] / 13
    print(f"{name}: {sum(ts)/len(ts):.4f} ms  n={p.shape[0]} m={mm} {per}", flush=True)
    t.close()


run("base", pos, attrs)
run("base, first half of the samples", pos, attrs, q[: m // 2].contiguous())
ne = 95435
ep, ea = pkg.synth.cloud_device(ne, w.seed + 77, u0=L, u1=L + 1.9)
run("strip outside x in [1000,1001.9]", torch.cat([pos, ep]).contiguous(), torch.cat([attrs, ea]).contiguous())
ep2, ea2 = pkg.synth.cloud_device(ne, w.seed + 78, u0=500.0, u1=501.9)
run("strip inside  x in [500,501.9]", torch.cat([pos, ep2]).contiguous(), torch.cat([attrs, ea2]).contiguous())
ep3, ea3 = pkg.synth.cloud_device(ne, w.seed + 79)
run("extras uniform inside", torch.cat([pos, ep3]).contiguous(), torch.cat([attrs, ea3]).contiguous())
run("first 49.9M", pos[:49900000].contiguous(), attrs[:49900000].contiguous())
run("first 49M", pos[:49000000].contiguous(), attrs[:49000000].contiguous())
