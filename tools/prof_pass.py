"""Where does a pass's time go beyond the main kernel?  Times one pass with different L2 states
(write-flush as in bench.py, read-flush, none) and with events around the library call only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge

pkg = ge.package()
w = pkg.synth.CONFIGS["cfg2"]
torch.cuda.set_device(0)
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv)
m, k = q.shape[0], w.k
tree = pkg.DeviceTree(pos, attrs)
idx = torch.empty((m, k), dtype=torch.int32, device="cuda")
rgba = torch.empty((m, 4), dtype=torch.uint8, device="cuda")
nrm = torch.empty((m, 3), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
big = torch.empty(64 << 20, dtype=torch.float32, device="cuda").zero_()


def run(mode, variant, reps=9):
    pkg.set_option("knn_variant", variant)
    ts = []
    for _ in range(reps):
        if mode == "write":
            flush.zero_()
        elif mode == "read":
            big.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tree.query(q, k, idx=idx, rgba=rgba, normal=nrm)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for variant in (6, 5, 2):
    print(f"variant {variant}: " + "  ".join(f"{mode} {run(mode, variant):.3f} ms" for mode in ("write", "read", "none")), flush=True)
