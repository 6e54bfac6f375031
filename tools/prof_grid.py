"""Grid kernel (variant 6) against the thread kernel (variant 2) on one workload: bit-identical
outputs, time per pass (L2 flushed), hand-over counts.  usage: prof_grid.py [cfg] [n] [grid] [k]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge

pkg = ge.package()
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = pkg.synth.CONFIGS[cfg]
n = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) else w.n_points
g = int(sys.argv[3]) if len(sys.argv) > 3 and int(sys.argv[3]) else w.gu
k = int(sys.argv[4]) if len(sys.argv) > 4 and int(sys.argv[4]) else w.k
torch.cuda.set_device(0)
pos, attrs = pkg.synth.cloud_device(n, w.seed, kind=w.kind, sigma=w.sigma)
q = pkg.synth.samples_device(g, g, center=w.center)
m = q.shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def outputs():
    return (torch.empty((m, k), dtype=torch.int32, device="cuda"), torch.empty((m, k), dtype=torch.float64, device="cuda"),
            torch.empty((m, 4), dtype=torch.uint8, device="cuda"), torch.empty((m, 3), dtype=torch.float32, device="cuda"))


def timed(tree, outs, want_d2, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tree.query(q, k, radius=w.radius, idx=outs[0], d2=outs[1] if want_d2 else None, rgba=outs[2], normal=outs[3])
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


results = {}
MODES = (("thread/order2", 2, 2, 1), ("scan/order2", 2, 5, 1), ("thread/order1", 1, 2, 1),
         ("grid/tma", 1, 6, 1), ("grid/ldgsts", 1, 6, 0))
if os.environ.get("PROF_GRID_SHORT"):
    MODES = (MODES[0], MODES[3], MODES[4])
for name, order, variant, tma in MODES:
    pkg.set_option("order", order)
    pkg.set_option("knn_variant", variant)
    pkg.set_option("grid_tma", tma)
    t0 = time.perf_counter()
    tree = pkg.DeviceTree(pos, attrs)
    info = tree.info()
    outs = outputs()
    tree.query(q, k, radius=w.radius, idx=outs[0], d2=outs[1], rgba=outs[2], normal=outs[3])
    torch.cuda.synchronize()
    fb = tree.fallback_counts()
    med, best = timed(tree, outs, False)
    results[name] = [o.clone() for o in outs]
    print(f"{name:14s} build {info.build_ms:8.2f} ms  pass median {med:7.3f} ms best {best:7.3f} ms "
          f"({m / med / 1e3:7.1f} M samples/s)  warp-fallback {fb[0]}  grid-handover {fb[1]}  "
          f"index {info.device_bytes / 1e6:.0f} MB", flush=True)
    tree.close()
ref = results["thread/order2"]
for name, outs in results.items():
    same = [bool((a == b).all()) for a, b in zip(outs, ref)]
    print(f"{name:14s} idx/d2/rgba/normal identical to thread/order2: {same}")
    if not all(same[:3]):
        bad = (outs[0] != ref[0]).any(dim=1).nonzero().view(-1)
        print("   first differing samples:", bad[:8].tolist(), "of", int(bad.numel()))
        if bad.numel():
            s = int(bad[0])
            print("   got", outs[0][s].tolist(), "\n   ref", ref[0][s].tolist())
            print("   got d2", outs[1][s].tolist(), "\n   ref d2", ref[1][s].tolist())
