"""Diagnosis of the multi-GPU step time: run with torchrun --nproc-per-node 2.  Times, per rank,
the plain query on (a) the slab index without ids, (b) the ghost-augmented index with ids, and
(c) the SlabTransfer step, before and after the NCCL communicator exists."""
import os, sys, math; sys.path.insert(0, "/root/repo")
import torch, torch.distributed as dist, __graft_entry__ as ge
pkg = ge.package()
if os.environ.get("PT_VARIANT"): pkg.set_option("knn_variant", int(os.environ["PT_VARIANT"]))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
w = pkg.synth.CONFIGS["cfg2"]; L = pkg.synth.L_DOMAIN; n = w.n_points; k = w.k
pos, attrs = pkg.synth.cloud_device(n, w.seed, u0=rank * L, u1=(rank + 1) * L, first_index=rank * n, device=dev)
q = pkg.synth.samples_device(w.gu, w.gv, u0=rank * L, u1=(rank + 1) * L, device=dev); m = q.shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
idx = torch.empty((m, k), dtype=torch.int32, device=dev); rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
nrm = torch.empty((m, 3), dtype=torch.float32, device=dev); d2 = torch.empty((m, k), dtype=torch.float64, device=dev)


def timeit(name, fn, n_it=13):
    ts = []
    for it in range(n_it):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if it >= 3: ts.append(a.elapsed_time(b))
    print(f"[rank {rank}] {name}: {sum(ts)/len(ts):.4f} ms (min {min(ts):.4f})", flush=True)


t0 = pkg.DeviceTree(pos, attrs)
timeit("plain index, before NCCL init", lambda: t0.query(q, k, idx=idx, rgba=rgba, normal=nrm))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
timeit("plain index, after NCCL init", lambda: t0.query(q, k, idx=idx, rgba=rgba, normal=nrm))
t0.close()
if world > 1:
    rk = math.sqrt(k / (math.pi * (n / (L * L)))); halo = 6.0 * rk
    ids = torch.arange(rank * n, (rank + 1) * n, dtype=torch.int32, device=dev)
    own_box = pkg.dist.points_box(pos); boxes = pkg.dist.gather_boxes(own_box)
    t1 = pkg.DeviceTree(pos, attrs, ids)
    timeit("own points + ids", lambda: t1.query(q, k, idx=idx, d2=d2, rgba=rgba, normal=nrm))
    t1.close()
    gpos, gattrs, gids = pkg.dist.exchange_ghosts(pos, attrs, ids, boxes, halo)
    t2 = pkg.DeviceTree(gpos, gattrs, gids)
    timeit("ghost-augmented + ids", lambda: t2.query(q, k, idx=idx, d2=d2, rgba=rgba, normal=nrm))
    slab = pkg.dist.SlabTransfer(pkg.dist.CudaSlabEngine(t2), own_box=own_box, halo=halo)
    timeit("slab.transfer validate=False", lambda: slab.transfer(q, k, validate=False))
    print(rank, "validate", slab.validate(), slab.stats, flush=True)
    dist.barrier(); dist.destroy_process_group()
