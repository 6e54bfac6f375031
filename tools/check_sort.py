"""The hand-written radix sort against cub::DeviceRadixSort: both are stable LSD sorts of the
same keys, so the two indexes must be identical -- checked through the query results and, on a
-DPT_STATS build, the traversal work counters -- and the build times are printed."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
w = pkg.synth.CONFIGS["cfg2"]; k = w.k
n = int(sys.argv[1]) if len(sys.argv) > 1 else w.n_points
pos, attrs = pkg.synth.cloud_device(n, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv); m = q.shape[0]
res = {}
for sort in (1, 0):
    pkg.set_option("sort", sort)
    pkg.DeviceTree(pos, attrs).close()
    t = pkg.DeviceTree(pos, attrs)
    idx = torch.empty((m, k), dtype=torch.int32, device=dev); d2 = torch.empty((m, k), dtype=torch.float64, device=dev)
    pkg.api.debug_stats()
    t.query(q, k, idx=idx, d2=d2); torch.cuda.synchronize()
    st = pkg.api.debug_stats()
    res[sort] = (idx, d2, st)
    print(f"sort={sort}: build {t.info().build_ms:.2f} ms  counters {dict(list(st.items())[:4]) if st['samples'] else '(no stats build)'}", flush=True)
    t.close()
pkg.set_option("sort", 1)
ok = bool(torch.equal(res[0][0], res[1][0])) and bool(torch.equal(res[0][1], res[1][1])) and res[0][2] == res[1][2]
print("indexes identical:", ok)
sys.exit(0 if ok else 1)
