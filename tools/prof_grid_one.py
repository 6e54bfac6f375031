"""One workload through the grid kernel only (for ncu and for the -DPT_STATS build).
usage: prof_grid_one.py [cfg] [n] [grid] [k] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge

pkg = ge.package()
a = sys.argv[1:] + [""] * 5
cfg = a[0] or "cfg2"
w = pkg.synth.CONFIGS[cfg]
n = int(a[1] or 0) or w.n_points
g = int(a[2] or 0) or w.gu
k = int(a[3] or 0) or w.k
reps = int(a[4] or 3)
torch.cuda.set_device(0)
if os.environ.get("PT_L2_GRAN"):      # experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes)
    import ctypes
    torch.zeros(1, device="cuda")
    rt = ctypes.CDLL("libcudart.so.12")
    v = ctypes.c_size_t(0)
    print("set", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["PT_L2_GRAN"]))), "get", rt.cudaDeviceGetLimit(ctypes.byref(v), 5), v.value)
for o in ("grid_min_occ10", "grid_lookup_cost", "grid_admit100"):
    if os.environ.get("PT_" + o.upper()):
        pkg.set_option(o, int(os.environ["PT_" + o.upper()]))
pkg.set_option("verbose", 1)
slab = int(os.environ.get("PT_SLAB", "1"))       # PT_SLAB=8: one of 8 x-slabs of the workload (n = its points)
u1 = pkg.synth.L_DOMAIN / slab
pos, attrs = pkg.synth.cloud_device(n, w.seed, u0=0.0, u1=u1, kind=w.kind, sigma=w.sigma)
q = pkg.synth.samples_device(g // slab, g, u0=0.0, u1=u1, center=w.center)
m = q.shape[0]
tree = pkg.DeviceTree(pos, attrs)
pkg.set_option("verbose", 0)
if os.environ.get("PT_TMA"):
    pkg.set_option("grid_tma", int(os.environ["PT_TMA"]))
if os.environ.get("PT_PAIR"):
    pkg.set_option("grid_pair", int(os.environ["PT_PAIR"]))
if os.environ.get("PT_VARIANT"):
    pkg.set_option("knn_variant", int(os.environ["PT_VARIANT"]))
idx = torch.empty((m, k), dtype=torch.int32, device="cuda")
rgba = torch.empty((m, 4), dtype=torch.uint8, device="cuda")
nrm = torch.empty((m, 3), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pkg.api.debug_stats(reset=True)
pkg.set_option("verbose", 1)
for r in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tree.query(q, k, radius=w.radius, idx=idx, rgba=rgba, normal=nrm)
    e1.record()
    torch.cuda.synchronize()
    pkg.set_option("verbose", 0)
    print(f"pass {r}: {e0.elapsed_time(e1):.3f} ms, fallback counts {tree.fallback_counts()}")
st = pkg.api.debug_stats(reset=False)
if st["grid_attempts"]:
    s = m * reps
    print({kk: round(v / s, 3) for kk, v in st.items() if v})
