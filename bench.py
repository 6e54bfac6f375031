#!/usr/bin/env python
"""bench.py -- mesh-sample kNN transfers/sec on B200 (BASELINE.json metric).

One "step" = one pass of the hot path (exact k-NN + fused colour/normal blend of every mesh
sample against the resident spatial index) over one batch of synthetic samples.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

`value`   : samples/s with inputs resident in HBM (CUDA events around the kernels, L2 flushed
            between steps), whole job = sum over ranks / max-over-ranks time.
`e2e`     : same metric through the reference-facing host C ABI (pt_transfer) with pinned HOST
            buffers: H2D of the 80-byte Point query records + kernels + D2H of the results
            inside the timed region.
`roofline`: algorithmic bytes (32 + 36k per sample) / kernel time vs the measured
            HBM copy peak (MEASURED_PEAKS.json).
`cpu_baseline`: the oracle's CGAL-style kd-tree (a port; the reference needs CGAL which is
            absent) on the host cores over a bounded window of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mesh-sample kNN transfers/sec"
UNIT = "samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--points", type=int, default=0, help="override points per GPU")
    ap.add_argument("--grid", type=int, default=0, help="override sample grid side per GPU")
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--variant", type=int, default=-1, help="knn kernel variant (tuning)")
    ap.add_argument("--order", type=int, default=-1, help="0 Morton, 1 Hilbert (tuning)")
    ap.add_argument("--sort", type=int, default=-1, help="1 hand-written radix sort, 0 CUB (tuning)")
    ap.add_argument("--slab", type=int, default=-1, help="(diagnosis) use the slab of this rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-window", type=float, default=0.25,
                    help="side fraction of the domain used for the bounded CPU sample")
    return ap.parse_args()


def algorithmic_bytes_per_sample(k):
    # 16 (query) + 16k (winner positions) + 16k (winner attrs) + 4k (idx out)
    # + 16 (blended rgba8 + normal out)
    return 32 + 36 * k


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the query kernel from the
# `ncu --set full` captures committed under profiles/ (r1_knn_{scan,thread}_kernel_ncu_summary.txt):
# valid only for the workload / kernel it was captured on.
NCU_TRAFFIC = {("cfg2", 16, 5): 1011.36e6 + 16.67e6, ("cfg2", 16, 2): 813.81e6 + 18.40e6}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled
    from a thread every ~0.5 ms (the region is tens of milliseconds, too short for nvidia-smi's
    loop mode, which is only the fallback)."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x4, "sw_power_cap"))

    def __init__(self, index, pci_bus_id=None):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self.stop_flag, self.thread, self.proc, self.rows = False, None, None, []
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if pci_bus_id:
                try:
                    h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode())
                except Exception:
                    h = None
            self.handle = h or pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        n, h = self.nvml, self.handle
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(int(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                self.mask |= int(reasons(h))
            except Exception:
                pass
            time.sleep(0.0005)

    def start(self):
        if self.nvml:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(name for bit, name in self.REASONS if self.mask & bit),
                    "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4)
                          if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


def resolve_workload(args, pkg):
    w = pkg.synth.CONFIGS[args.workload]
    n = args.points or w.n_points
    gu = args.grid or w.gu
    gv = args.grid or w.gv
    k = args.k or w.k
    return w, n, gu, gv, k


# ---------------------------------------------------------------------------------------------
def cpu_window(pkg, pto, pos, attrs, w, frac, L0, L1, n_samples=100_000):
    """Bounded CPU sample: a frac x frac window of the slab -- the cloud points inside it (plus
    a margin so every neighbour is present) and n_samples samples on the same surface."""
    import torch
    side = (L1 - L0) * frac
    margin = 2.0
    u0, v0 = L0, 0.0
    pm = ((pos[:, 0] >= u0 - margin) & (pos[:, 0] < u0 + side + margin) &
          (pos[:, 1] >= v0 - margin) & (pos[:, 1] < v0 + side + margin))
    sub_pos = pos[pm].contiguous()
    sub_attr = attrs[pm].contiguous()
    g = max(2, int(round(n_samples ** 0.5)))
    sub_q = pkg.synth.samples_device(g, g, u0=u0, u1=u0 + side, v0=v0, v1=v0 + side,
                                     center=True, device=pos.device)
    P = pkg.synth.points_to_host(sub_pos, sub_attr)
    Q = pkg.synth.queries_to_host(sub_q)
    return P, Q


def run_cpu_leg(pkg, pto, P, Q, k, radius, steps, warmup):
    threads = pto.max_threads()
    t0 = time.perf_counter()
    tree = pto.KdTree(P)
    build_s = time.perf_counter() - t0
    r = -1.0 if radius is None else radius
    for _ in range(max(1, warmup)):
        tree.knn(Q[: max(1, len(Q) // 8)], k, radius=r, exact_ties=False, want_d2=False)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        idx, d2 = tree.knn(Q, k, radius=r, exact_ties=False)
        pto.blend(P, idx, d2)
        times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    # SURVEY 8 M4 also asks for the single-thread rate (the reference is single-threaded unless
    # built with -DMULTI_THREADING=ON): the k-NN alone, one thread, on an eighth of the sample
    Q1 = Q[: max(1, len(Q) // 8)]
    t0 = time.perf_counter()
    tree.knn(Q1, k, radius=r, exact_ties=False, nthreads=1, want_d2=False)
    t1 = time.perf_counter() - t0
    tree.close()
    return {"value": len(Q) / t, "unit": UNIT, "cores": threads, "kind": "port",
            "single_thread_knn_value": len(Q1) / t1,
            "sample": f"{len(Q)} samples (unique-vertex queries, k={k}) against a {len(P)}-point "
                      f"window of the workload cloud; CGAL-style kd-tree (sliding midpoint, bucket "
                      f"10) + blend, OpenMP over samples; tree build {build_s:.2f} s excluded",
            "build_s": build_s, "ms_per_step": t * 1e3}


# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm must use all the
        # host cores it can (rank 0 runs alone), so set it before any OpenMP runtime loads
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference_arm(args, pkg, world, rank, local_rank)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.variant >= 0:
        pkg.set_option("knn_variant", args.variant)
    if args.order >= 0:
        pkg.set_option("order", args.order)
    if args.sort >= 0:
        pkg.set_option("sort", args.sort)

    w, n, gu, gv, k = resolve_workload(args, pkg)
    L = pkg.synth.L_DOMAIN
    slab_id = rank if args.slab < 0 else args.slab
    u0, u1 = slab_id * L, (slab_id + 1) * L      # weak scaling: one slab of the scan per rank
    pos, attrs = pkg.synth.cloud_device(n, w.seed, u0=u0, u1=u1, kind=w.kind, sigma=w.sigma,
                                        first_index=rank * n, device=dev)
    q = pkg.synth.samples_device(gu, gv, u0=u0, u1=u1, center=w.center, device=dev)
    m = q.shape[0]
    torch.cuda.synchronize()
    ids = own_box = halo = None
    n_own = n
    if world > 1:
        # slab + ghost zone: the points of the other slabs within `halo` of this slab's box are
        # exchanged ONCE here, so the steady-state step needs no collective
        import math
        rk = math.sqrt(k / (math.pi * (n / (L * L))))       # expected k-th neighbour distance
        halo = max(6.0 * rk, 2.0 * (w.radius or 0.0))
        ids = torch.arange(rank * n, (rank + 1) * n, dtype=torch.int32, device=dev)
        own_box = pkg.dist.points_box(pos)
        boxes = pkg.dist.gather_boxes(own_box)
        pos, attrs, ids = pkg.dist.exchange_ghosts(pos, attrs, ids, boxes, halo)
        n = pos.shape[0]
    pkg.DeviceTree(pos, attrs, ids).close()     # warm-up build (module load, allocator)
    t0 = time.perf_counter()
    tree = pkg.DeviceTree(pos, attrs, ids)
    build_wall_ms = (time.perf_counter() - t0) * 1e3
    info = tree.info()

    idx = torch.empty((m, k), dtype=torch.int32, device=dev)
    rgba = torch.empty((m, 4), dtype=torch.uint8, device=dev)
    nrm = torch.empty((m, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    slab = None
    if world > 1:
        # slab-sharded: owner k-NN, halo exchange of the boundary samples over NCCL, K5 merge
        slab = pkg.dist.SlabTransfer(pkg.dist.CudaSlabEngine(tree), own_box=own_box, halo=halo)
    result = {}

    def step():
        if slab is None:
            tree.query(q, k, radius=w.radius, idx=idx, rgba=rgba, normal=nrm)
        else:
            result.update(slab.transfer(q, k, radius=w.radius, validate=False))

    for _ in range(max(3, args.warmup)):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:
        bus = None
    sampler = ClockSampler(local_rank, bus)
    sampler.start()
    launches0 = pkg.kernel_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    torch.cuda.synchronize()
    host_t0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()
        a.record()
        step()
        b.record()
    host_issue_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps   # CPU time to enqueue a step
    torch.cuda.synchronize()
    launches = pkg.kernel_launch_count() - launches0
    if slab is not None and not slab.validate():
        raise SystemExit("halo capacity overflowed during the timed steps: results invalid")
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    ms_per_rank = [dev_ms / args.steps]
    if world > 1:
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        ms_per_rank = [float(x.item()) / args.steps for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * m / (ms_per_step * 1e-3)

    # ---- e2e through the host C ABI with pinned host buffers --------------------------------
    q_host = pkg.synth.queries_to_host(q, pinned=True)                  # [m,80] uint8 pinned
    out_idx = torch.empty((m, k), dtype=torch.int32, pin_memory=True)
    out_rgba = torch.empty((m, 4), dtype=torch.uint8, pin_memory=True)
    out_nrm = torch.empty((m, 3), dtype=torch.float32, pin_memory=True)
    qh_np = q_host.numpy().view(pkg.POINT_DTYPE).reshape(-1)
    out = {"idx": out_idx.numpy(), "rgba": out_rgba.numpy(), "normal": out_nrm.numpy()}
    q_xyz_host = torch.empty((m, 3), dtype=torch.float64, pin_memory=True)
    q_xyz_host.copy_(q)

    def e2e_step():
        if slab is None:
            tree.transfer(qh_np, k, radius=w.radius, out=out)      # pt_transfer: host in, host out
        else:
            # multi-GPU public API: pinned host samples in, results in pinned host memory
            # (H2D, owner step and D2H pipelined in chunks; one deferred validation)
            slab.transfer_host(q_xyz_host, k, {"idx": out_idx, "rgba": out_rgba, "normal": out_nrm},
                               radius=w.radius)

    for _ in range(3):
        e2e_step()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    info2 = tree.info()
    ref_idx = idx if slab is None else result["idx"]
    assert bool((out_idx.to(dev) == ref_idx).all()), "host-buffer and device-buffer paths disagree"

    variant_used = pkg.get_option("knn_variant")
    if variant_used < 0:      # auto rule of launch_query (pt_knn.cu)
        variant_used = 0 if m <= 12288 else (
            5 if (k > 16 or (w.radius is None and m >= 100000 and slab is None)) else 2)
    peak, peak_src = measured_peaks()
    alg_bytes = algorithmic_bytes_per_sample(k) * m
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w.name, "points_per_gpu": n_own, "ghost_points_per_gpu": n - n_own,
                   "samples_per_gpu": m, "k": k,
                   "radius": w.radius, "coord_storage": "f32x4" if info.coord_mode == 1 else "f64",
                   "l2": "flushed between steps (256 MiB write)",
                   "parallelism": f"slab x{world}", "knn_variant": variant_used,
                   "order": pkg.get_option("order")},
        "e2e": {"value": world * m / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": m * (80 if slab is None else 24),
                "d2h_bytes_per_step": m * (4 * k + 4 + 12), "ms_per_step": e2e_s * 1e3,
                "h2d_ms": info2.last_h2d_ms, "kernel_ms": info2.last_query_ms,
                "d2h_ms": info2.last_d2h_ms},
        "gpu_launches": int(launches), "ms_per_rank": ms_per_rank,
        "host_issue_ms_per_step": host_issue_ms,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (NCU_TRAFFIC.get((args.workload, k, variant_used))
                                 if world == 1 and not (args.points or args.grid) else None),
                     "traffic_source": "ncu --set full, profiles/r1_knn_scan_kernel_ncu_summary.txt",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel": "knn_*_kernel",
                     "kernel_ms": ms_per_step},
        "build": {"ms": info.build_ms, "wall_ms": build_wall_ms, "points_per_s": n / (info.build_ms * 1e-3),
                  "halo": halo,
                  "leaves": int(info.n_leaves), "index_bytes": int(info.device_bytes)},
    }
    if slab is not None:
        line["exchange"] = dict(slab.stats, samples=m, crossing_fast=slab.crossing_count())
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pto
        P, Q = cpu_window(pkg, pto, pos, attrs, w, args.cpu_window, u0, u1)
        line["cpu_baseline"] = run_cpu_leg(pkg, pto, P, Q, k, w.radius, steps=3, warmup=1)
    if rank == 0:
        print(json.dumps(line))
    tree.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference_arm(args, pkg, world, rank, local_rank):
    """The reference's own CPU implementation of the path (kd-tree k-NN per sample), all host
    threads, on the same workload: the CGAL-linked original cannot be built here, so this is
    the oracle port (cpu_baseline.kind = "port").  Rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import pto
    w, n, gu, gv, k = resolve_workload(args, pkg)
    L = pkg.synth.L_DOMAIN
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        # same generator as the GPU arm; only a bounded window is copied to the host
        side = L * args.cpu_window
        pos, attrs = pkg.synth.cloud_device(n, w.seed, kind=w.kind, sigma=w.sigma)
        P, Q = cpu_window(pkg, pto, pos, attrs, w, args.cpu_window, 0.0, L)
        del pos, attrs
        torch.cuda.empty_cache()
    else:
        side = 100.0
        P = pkg.synth.cloud_host(int(n * (side / L) ** 2), w.seed, side=side)
        Q = pkg.synth.samples_host(316, side=side)
    leg = run_cpu_leg(pkg, pto, P, Q, k, w.radius, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": leg["value"], "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w.name, "points_per_gpu": n, "samples_per_gpu": gu * gv, "k": k,
                   "radius": w.radius},
        "cpu_baseline": leg,
        "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
