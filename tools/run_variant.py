"""Runs the cfg2 query a few times with the given kernel variant (for ncu captures)."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w = pkg.synth.CONFIGS["cfg2"]; k = w.k
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed)
if len(sys.argv) > 2:       # optional: use only the first n points (size-sensitivity captures)
    nn = int(sys.argv[2]); pos, attrs = pos[:nn].contiguous(), attrs[:nn].contiguous()
q = pkg.synth.samples_device(w.gu, w.gv); m = q.shape[0]
pkg.set_option("knn_variant", variant)
tree = pkg.DeviceTree(pos, attrs)
idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
nrm = torch.empty((m, 3), dtype=torch.float32, device=dev)
for _ in range(3):
    tree.query(q, k, idx=idx, rgba=rgba, normal=nrm)
torch.cuda.synchronize(); tree.close()
