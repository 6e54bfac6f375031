import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, torch.distributed as dist
import __graft_entry__ as ge
pkg = ge.package()
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
w = pkg.synth.CONFIGS["cfg2"]; L = 1000.0
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed, u0=rank*L, u1=(rank+1)*L, first_index=rank*w.n_points, device=dev)
q = pkg.synth.samples_device(w.gu, w.gv, u0=rank*L, u1=(rank+1)*L, device=dev)
tree = pkg.DeviceTree(pos, attrs)
eng = pkg.dist.CudaSlabEngine(tree); st = pkg.dist.SlabTransfer(eng)
k = w.k; R = world; cap = st.cap
names = ["own_query", "route", "a2a_q", "prepare", "halo_query", "a2a_back", "merge"]
acc = {n: 0.0 for n in names}
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(23):
    e = [ev()]
    own, out = eng.query(q, k, outputs=True); e.append(ev())
    h = eng.halo_buffers(R, cap, k); h["flag"].zero_()
    eng.halo_route(q, own, k, None, st.boxes6, rank, cap, h); e.append(ev())
    dist.all_to_all_single(h["recv"].view(R, -1), h["send"].view(R, -1)); e.append(ev())
    eng.halo_prepare(h, R, cap); e.append(ev())
    eng.tree.query(h["hq"], k, radius2_per_query=h["hr2"], cand=h["hcand"].view(-1)); e.append(ev())
    dist.all_to_all_single(h["back"].view(R, -1), h["hcand"].view(R, -1)); e.append(ev())
    for r in range(R):
        if r != rank: eng.halo_merge(own, h, r, cap, k, out)
    e.append(ev())
    torch.cuda.synchronize()
    if it >= 3:
        for i, n in enumerate(names): acc[n] += e[i].elapsed_time(e[i+1])
if rank == 0:
    print({n: round(v / 20, 4) for n, v in acc.items()}, "total", round(sum(acc.values())/20, 4))
dist.destroy_process_group()
