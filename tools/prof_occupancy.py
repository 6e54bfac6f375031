"""Occupancy sensitivity of the query kernel: pads the dynamic shared memory of every block
(pt_set_option smem_pad) so fewer one-warp blocks fit per SM, and times cfg2."""
import sys; sys.path.insert(0, "/root/repo")
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
w = pkg.synth.CONFIGS["cfg2"]
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed)
q = pkg.synth.samples_device(w.gu, w.gv); m, k = q.shape[0], w.k
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tree = pkg.DeviceTree(pos, attrs)
idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
nrm = torch.empty((m, 3), dtype=torch.float32, device=dev)
for pad in (0, 2048, 4096, 8192, 16384, 32768, 65536):
    pkg.set_option("smem_pad", pad)
    ts = []
    for it in range(13):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); tree.query(q, k, idx=idx, rgba=rgba, normal=nrm); b.record(); torch.cuda.synchronize()
        if it >= 3: ts.append(a.elapsed_time(b))
    blk = 32 * ((k + 8) * 12 + 24 * 8) + 1024 + pad
    print(f"pad={pad:6d}  blocks/SM~{min(32, (228 * 1024) // blk):3d}  {sum(ts)/len(ts):.4f} ms")
pkg.set_option("smem_pad", 0)
tree.close()
