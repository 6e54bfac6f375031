"""Detail transfer onto a cloud sharded across the GPUs of one box, samples arriving UNSORTED.

north_star: "Each GPU indexes a spatial slab of the cloud plus a halo sized to the search radius.
Queries are broadcast, or routed by slab, and the per-rank top-k candidates are merged over NCCL".
The reference has no distributed path (its query loop is src/pointsTransfer.cpp:465-479); this is
the multi-GPU form of that loop.  One process per GPU (``torch.distributed`` / NCCL):

  1. route    every rank holds an arbitrary part of the samples; ``pt_route_samples_device``
              sorts them into fixed-capacity blocks per owning slab (x-range cuts), one
              ``all_to_all`` delivers them                                   -- no host sync
  2. owner    the slab step of ``dist.SlabTransfer`` on the received rows: k-NN + blend on the
              ghost-augmented index with the ghost-zone check (no collective), or -- without a
              ghost zone -- halo route / bounded halo search / K5 merge over two more all_to_all
  3. return   neighbour ids, colours and normals travel back (``all_to_all``) and
              ``pt_scatter_rows_device`` puts them in the caller's sample order.

Capacity overflows and ghost-zone violations are accumulated on the device and checked once per
batch of steps (``validate``), so a step never waits for the host.
"""
import torch
import torch.distributed as dist

from . import dist as slabs


class ShardedTransfer:
    def __init__(self, engine, cuts, k, m_local, own_box=None, halo=None, group=None, slack=1.3):
        """engine: this rank's slab engine (``dist.CudaSlabEngine(tree)``; index ghost-augmented
        iff ``halo``); cuts: R+1 x-cuts (first -inf, last +inf); m_local: samples a call may
        bring on this rank."""
        self.group = group
        self.R = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dev = engine.device
        self.k = int(k)
        self.cuts = torch.as_tensor(cuts, dtype=torch.float64).to(self.dev).contiguous()
        assert self.cuts.numel() == self.R + 1
        self.slab = slabs.SlabTransfer(engine, group=group, own_box=own_box, halo=halo)
        self.cap = ((int(slack * m_local / self.R) + 1024 + 31) // 32) * 32
        R, cap, k, dev = self.R, self.cap, self.k, self.dev
        self.route = None
        self.recv = torch.empty((R, cap, 4), dtype=torch.float64, device=dev)
        self.ret = {"idx": torch.empty((R * cap, k), dtype=torch.int32, device=dev),
                    "rgba": torch.empty((R * cap, 4), dtype=torch.uint8, device=dev),
                    "normal": torch.empty((R * cap, 3), dtype=torch.float32, device=dev)}
        self.flag_acc = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.last_counts = None

    def transfer(self, q, out, radius=None):
        """q float64 [m,3] on this rank's device (any samples, any order); out: dict of device
        tensors idx [m,k] int32, rgba [m,4] uint8, normal [m,3] float32 (filled in sample order).
        Asynchronous; call ``validate`` before trusting a batch of steps."""
        R, cap, k = self.R, self.cap, self.k
        self.route = slabs.route_samples(q, self.cuts, cap, out=self.route)
        send, sel, counts, overflow = self.route
        torch.maximum(self.flag_acc, overflow, out=self.flag_acc)
        dist.all_to_all_single(self.recv.view(R, -1), send.view(R, -1), group=self.group)
        rows = self.recv.view(R * cap, 4)
        rq = rows[:, :3].contiguous()
        rb = rows[:, 3].contiguous()                # +inf (a sample) or NaN (unused row)
        fast = getattr(self.slab.engine, "fast", False)
        res = self.slab.transfer(rq, k, radius=radius, validate=not fast, r2pq=rb)
        for name in ("idx", "rgba", "normal"):
            dist.all_to_all_single(self.ret[name].view(R, -1), res[name].view(R, -1), group=self.group)
            slabs.scatter_rows(self.ret[name], sel.view(-1), out[name])
        self.last_counts = counts
        return out

    def transfer_host(self, q_host, out_host, radius=None, scratch=None, pieces=1):
        """Host-buffer form: q_host float64 [m,3] (pinned), out_host dict of (pinned) host tensors.
        With ``pieces`` > 1 the batch runs as consecutive sharded steps so that the D2H copy of one
        piece's results overlaps the next piece's step (every rank must use the same ``pieces``);
        measured on 4 GPUs at cfg3 this LOSES (2.36 ms against 1.67 ms in one piece: every step
        pays its four all_to_all and the small-launch inefficiency of the query kernel), hence the
        default of one piece.  Returns after the results have landed."""
        m = q_host.shape[0]
        cuda = self.dev.type == "cuda"
        if scratch is None or scratch["q"].shape[0] != m:
            scratch = {"q": torch.empty((m, 3), dtype=torch.float64, device=self.dev),
                       "idx": torch.empty((m, self.k), dtype=torch.int32, device=self.dev),
                       "rgba": torch.empty((m, 4), dtype=torch.uint8, device=self.dev),
                       "normal": torch.empty((m, 3), dtype=torch.float32, device=self.dev)}
            if cuda:
                scratch["copy_stream"] = torch.cuda.Stream(device=self.dev)
        pieces = max(1, min(int(pieces), max(m, 1)))
        step = -(-m // pieces)
        scratch["q"].copy_(q_host, non_blocking=True)
        cur = torch.cuda.current_stream(self.dev) if cuda else None
        for p in range(pieces):
            lo, hi = p * step, min(m, (p + 1) * step)
            part = {n: scratch[n][lo:hi] for n in ("idx", "rgba", "normal")}
            self.transfer(scratch["q"][lo:hi], part, radius=radius)
            if cuda:
                cs = scratch["copy_stream"]
                cs.wait_stream(cur)
                with torch.cuda.stream(cs):
                    for n in ("idx", "rgba", "normal"):
                        out_host[n][lo:hi].copy_(part[n], non_blocking=True)
            else:
                for n in ("idx", "rgba", "normal"):
                    out_host[n][lo:hi].copy_(part[n])
        if cuda:
            cur.wait_stream(scratch["copy_stream"])
            cur.synchronize()
        return scratch

    def validate(self):
        """True iff no step since the last call overflowed a routing / halo block or left the
        ghost zone (collective)."""
        ok_slab = self.slab.validate()
        dist.all_reduce(self.flag_acc, op=dist.ReduceOp.MAX, group=self.group)
        bad = int(self.flag_acc.item())
        self.flag_acc.zero_()
        return ok_slab and not bad

    def stats(self):
        """Routing / exchange counters of the last step (synchronises)."""
        routed = self.last_counts.to(torch.int64)
        away = int(routed.sum().item()) - int(routed[self.rank].item())
        s = dict(self.slab.stats)
        s["routed_to_other_slabs"] = away
        s["crossing"] = self.slab.crossing_count() if s.get("path") == "fast" else s.get("crossing", 0)
        if s.get("path") == "fast":
            s["halo_rows_per_peer"] = [int(v) for v in self.slab._last_counts.tolist()]
        return s
