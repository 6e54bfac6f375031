"""Slab-sharded detail transfer across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink).  Every rank owns one spatial
slab of the cloud and its own index; every sample is owned by exactly one rank.  A sample is
answered by its owner; only if the ball of its k-th neighbour reaches another slab's bounding
box does it take part in the single exchange step:

  1. owner:  local k-NN                          -> candidates + k-th squared distance (bound)
  2. route:  samples whose ball touches slab r   -> all_to_all (xyz + bound), usually a sliver
  3. remote: radius-bounded k-NN (d2 <= bound)   -> candidate lists, all_to_all back
  4. owner:  K5 merge kernel (key (d2, global id)) + blend of the merged list

so the result is bit-identical to a single index over the whole cloud, for any slab count.
The reference has no distributed path ("NCCL / MPI / Gloo: none"); this is
new plumbing around the same kernels.

The numerical work is delegated to an *engine* (``CudaSlabEngine`` in production).  The host
logic here only touches tensors through torch, so it also runs on CPU tensors over ``gloo``,
which is how the world_size-2 tests exercise it without a GPU.
"""
import torch
import torch.distributed as dist

CAND_BYTES = 32


def box_lower_bound2(q, lo, hi):
    """Squared distance from points q[m,3] to the box [lo, hi] (Distance.h:27-57), fp64, shrunk
    by a relative 1e-12 so that rounding can never exclude a slab that holds a neighbour."""
    e = torch.clamp(torch.maximum(lo - q, q - hi), min=0.0)
    return (e * e).sum(dim=1) * (1.0 - 1e-12)


class CudaSlabEngine:
    """The production engine: this rank's slab index on its GPU (hand-written CUDA kernels)."""

    def __init__(self, tree):
        from . import api
        self._api = api
        self.tree = tree
        info = tree.info()
        dev = tree.torch_device
        self.device = dev
        self.n_points = int(info.n_points)
        self.bbox_lo = torch.tensor(list(info.bbox_lo), dtype=torch.float64, device=dev)
        self.bbox_hi = torch.tensor(list(info.bbox_hi), dtype=torch.float64, device=dev)

    def query(self, q, k, radius=None, radius2_per_query=None, outputs=False, want_d2=False,
              want_cand=True):
        """-> (cand uint8 [m, k, 32] pt_cand records (d2, global id, attributes), out) where
        out is None or, with ``outputs``, the fused-blend results of the same launch."""
        m = q.shape[0]
        dev = q.device
        cand = torch.empty((m, k, CAND_BYTES), dtype=torch.uint8, device=dev) if want_cand else None
        out = None
        if outputs:
            out = {"idx": torch.empty((m, k), dtype=torch.int32, device=dev),
                   "rgba": torch.empty((m, 4), dtype=torch.uint8, device=dev),
                   "normal": torch.empty((m, 3), dtype=torch.float32, device=dev)}
            if want_d2:
                out["d2"] = torch.empty((m, k), dtype=torch.float64, device=dev)
        if m:
            self.tree.query(q, k, radius=radius, radius2_per_query=radius2_per_query,
                            cand=cand.view(-1) if want_cand else None,
                            idx=out["idx"] if out else None,
                            d2=out.get("d2") if out else None,
                            rgba=out["rgba"] if out else None,
                            normal=out["normal"] if out else None)
        return cand, out

    # ---- fixed-capacity halo exchange (no host synchronisation inside a step) --------------
    fast = True

    def halo_buffers(self, n_ranks, cap, k):
        key = (n_ranks, cap, k)
        if getattr(self, "_halo_key", None) != key:
            dev = self.device
            f64, u8, i32 = torch.float64, torch.uint8, torch.int32
            self._halo = {
                "send": torch.zeros((n_ranks, cap + 1, 4), dtype=f64, device=dev),
                "recv": torch.zeros((n_ranks, cap + 1, 4), dtype=f64, device=dev),
                "sel": torch.zeros((n_ranks, cap), dtype=i32, device=dev),
                "counts": torch.zeros((n_ranks,), dtype=i32, device=dev),
                "flag": torch.zeros((1,), dtype=i32, device=dev),
                "hq": torch.zeros((n_ranks * cap, 3), dtype=f64, device=dev),
                "hr2": torch.zeros((n_ranks * cap,), dtype=f64, device=dev),
                "hcand": torch.zeros((n_ranks * cap, k, CAND_BYTES), dtype=u8, device=dev),
                "back": torch.zeros((n_ranks, cap, k, CAND_BYTES), dtype=u8, device=dev),
            }
            self._halo_key = key
        return self._halo

    def ghost_check(self, q, d2, k, radius, boxes6, rank, halo, flag):
        api = self._api
        with torch.cuda.device(self.device):
            api._check(api.lib().pt_ghost_check_device(
                api._tptr(q), api._tptr(d2), q.shape[0], int(k), api._radius(radius),
                api._tptr(boxes6), boxes6.shape[0], int(rank), float(halo), api._tptr(flag),
                api._stream_ptr()), "pt_ghost_check_device")

    def halo_route(self, q, own, k, radius, boxes6, rank, cap, h):
        api = self._api
        with torch.cuda.device(self.device):
            api._check(api.lib().pt_halo_route_device(
                api._tptr(q), api._tptr(own), q.shape[0], int(k), api._radius(radius),
                api._tptr(boxes6), boxes6.shape[0], int(rank), int(cap), api._tptr(h["send"]),
                api._tptr(h["sel"]), api._tptr(h["counts"]), api._tptr(h["flag"]),
                api._stream_ptr()), "pt_halo_route_device")

    def halo_prepare(self, h, n_ranks, cap):
        api = self._api
        with torch.cuda.device(self.device):
            api._check(api.lib().pt_halo_prepare_device(
                api._tptr(h["recv"]), int(n_ranks), int(cap), api._tptr(h["hq"]),
                api._tptr(h["hr2"]), api._stream_ptr()), "pt_halo_prepare_device")

    def halo_merge(self, own, h, r, cap, k, out):
        api = self._api
        with torch.cuda.device(self.device):
            api._check(api.lib().pt_halo_merge_device(
                api._tptr(own), api._tptr(h["back"][r]), api._tptr(h["sel"][r]),
                api._tptr(h["counts"][r:r + 1]), int(cap), int(k), api._tptr(out["idx"]),
                api._tptr(out.get("d2")), api._tptr(out["rgba"]), api._tptr(out["normal"]),
                api._stream_ptr()), "pt_halo_merge_device")

    def merge(self, lists, k, want_d2=False):
        """lists uint8 [R, m, k, 32] -> dict(idx, rgba, normal[, d2])."""
        r, m = lists.shape[0], lists.shape[1]
        dev = lists.device
        out = {"idx": torch.empty((m, k), dtype=torch.int32, device=dev),
               "rgba": torch.empty((m, 4), dtype=torch.uint8, device=dev),
               "normal": torch.empty((m, 3), dtype=torch.float32, device=dev)}
        if want_d2:
            out["d2"] = torch.empty((m, k), dtype=torch.float64, device=dev)
        if m:
            self._api.merge_device(lists.contiguous().view(-1), r, m, k, idx=out["idx"],
                                   d2=out.get("d2"), rgba=out["rgba"], normal=out["normal"])
        return out


def route_samples(q, cuts, cap, out=None, stream=None):
    """pt_route_samples_device: q float64 [m,3] CUDA, cuts float64 [R+1] CUDA (-inf first, +inf
    last).  Returns (send [R,cap,4] f64, sel [R,cap] i32, counts [R] i32, overflow [1] i32);
    ``out`` = a previous result to reuse the buffers."""
    from . import api
    R = cuts.numel() - 1
    dev = q.device
    if not q.is_cuda:      # host-logic tests over gloo: the same contract in torch ops
        send = torch.full((R, cap, 4), float("nan"), dtype=torch.float64)
        sel = torch.full((R, cap), -1, dtype=torch.int32)
        counts = torch.zeros((R,), dtype=torch.int32)
        overflow = torch.zeros((1,), dtype=torch.int32)
        owner = torch.bucketize(q[:, 0].contiguous(), cuts[1:-1].contiguous(), right=True)
        for r in range(R):
            ids = torch.nonzero(owner == r).view(-1)
            counts[r] = ids.numel()
            if ids.numel() > cap:
                overflow[0] = 1
                ids = ids[:cap]
            send[r, :ids.numel(), :3] = q[ids]
            send[r, :ids.numel(), 3] = float("inf")
            sel[r, :ids.numel()] = ids.to(torch.int32)
        return send, sel, counts, overflow
    if out is None:
        out = (torch.empty((R, cap, 4), dtype=torch.float64, device=dev),
               torch.empty((R, cap), dtype=torch.int32, device=dev),
               torch.empty((R,), dtype=torch.int32, device=dev),
               torch.empty((1,), dtype=torch.int32, device=dev))
    send, sel, counts, overflow = out
    with torch.cuda.device(dev):
        api._check(api.lib().pt_route_samples_device(
            api._tptr(q), q.shape[0], api._tptr(cuts), R, int(cap), api._tptr(send), api._tptr(sel),
            api._tptr(counts), api._tptr(overflow), api._stream_ptr(stream)), "pt_route_samples_device")
    return out


def scatter_rows(src, sel, dst, stream=None):
    """pt_scatter_rows_device: dst[sel[t]] = src[t] for sel[t] >= 0 (row-major CUDA tensors)."""
    from . import api
    rows = sel.numel()
    if not src.is_cuda:
        valid = sel.view(-1) >= 0
        dst[sel.view(-1)[valid].long()] = src.view(rows, -1)[valid].view((-1,) + tuple(dst.shape[1:]))
        return
    row_bytes = src.element_size() * (src.numel() // max(rows, 1))
    assert src.is_contiguous() and dst.is_contiguous() and src.dtype == dst.dtype
    with torch.cuda.device(src.device):
        api._check(api.lib().pt_scatter_rows_device(api._tptr(src), api._tptr(sel), rows, row_bytes,
                                                    api._tptr(dst), api._stream_ptr(stream)),
                   "pt_scatter_rows_device")


def cand_d2(cand):
    """k-th ... view the squared distances of pt_cand records: cand uint8 [..., 32] -> f64 [...]."""
    return cand[..., :8].contiguous().view(torch.float64).squeeze(-1)


def empty_cand(shape, device):
    """Candidate records meaning "no neighbour": d2 = +inf, id = -1."""
    c = torch.zeros(tuple(shape) + (CAND_BYTES,), dtype=torch.uint8, device=device)
    c[..., :8] = torch.tensor([float("inf")], dtype=torch.float64).view(torch.uint8).to(device)
    c[..., 8:12] = torch.tensor([-1], dtype=torch.int32).view(torch.uint8).to(device)
    return c


def points_box(xyz):
    """[2,3] float64 bounding box of xyz[n,>=3] (inverted when empty)."""
    if xyz.shape[0] == 0:
        return torch.tensor([[float("inf")] * 3, [float("-inf")] * 3], dtype=torch.float64,
                            device=xyz.device)
    p = xyz[:, :3].to(torch.float64)
    return torch.stack([p.min(dim=0).values, p.max(dim=0).values])


def gather_boxes(box, group=None):
    """all_gather of every rank's [2,3] box -> [R,2,3]."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return box[None].clone()
    boxes = [torch.empty_like(box) for _ in range(dist.get_world_size(group))]
    dist.all_gather(boxes, box.contiguous(), group=group)
    return torch.stack(boxes)


def exchange_ghosts(pos, attrs, ids, boxes, halo, group=None):
    """Build-time ghost-zone exchange (north_star: "a spatial slab of the cloud plus a halo sized
    to the search radius").  Every rank sends each peer its own points that lie within ``halo``
    of that peer's slab box and receives the peers' points near its own box.  Returns the
    augmented (pos, attrs, ids), sorted by global id so that local tie-breaking stays global.
    pos [n,4], attrs [n,16] uint8, ids [n] int32 (strictly increasing)."""
    R = dist.get_world_size(group) if dist.is_initialized() else 1
    if R == 1:
        return pos, attrs, ids
    rank = dist.get_rank(group)
    dev = pos.device
    xyz = pos[:, :3].to(torch.float64)
    h2 = float(halo) * float(halo) * (1.0 + 1e-9)
    send_idx, counts = [], []
    for r in range(R):
        if r == rank or pos.shape[0] == 0:
            sel = torch.empty((0,), dtype=torch.int64, device=dev)
        else:
            e = torch.clamp(torch.maximum(boxes[r, 0] - xyz, xyz - boxes[r, 1]), min=0.0)
            sel = torch.nonzero((e * e).sum(dim=1) <= h2).view(-1)
        send_idx.append(sel)
        counts.append(int(sel.numel()))
    order = torch.cat(send_idx)
    sc = torch.tensor(counts, dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    rcl = rc.tolist()

    def a2a(t):
        out = torch.empty((sum(rcl),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_to_all_single(out, t[order].contiguous(), output_split_sizes=rcl,
                               input_split_sizes=counts, group=group)
        return out

    g_pos, g_attr, g_ids = a2a(pos), a2a(attrs), a2a(ids)
    all_ids = torch.cat([ids, g_ids])
    perm = torch.argsort(all_ids)
    return (torch.cat([pos, g_pos])[perm].contiguous(),
            torch.cat([attrs, g_attr])[perm].contiguous(), all_ids[perm].contiguous())


class SlabTransfer:
    """Collective detail transfer: every rank calls ``transfer`` with the samples it owns.

    With ``own_box`` and ``halo`` (index built over the slab plus the ghost zone returned by
    ``exchange_ghosts``) a sample is final as soon as ``dist(q, own_box) + r_k <= halo`` -- every
    point that could matter is local -- so the steady state needs no collective at all; samples
    that violate the bound make the step fall back to the per-step halo exchange."""

    def __init__(self, engine, group=None, own_box=None, halo=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        dev = engine.device
        self.halo = None if halo is None else float(halo)
        if halo is not None and own_box is None:
            # the engine's bbox already includes the ghost points: the ghost-zone test would be
            # evaluated against an inflated box and declare inexact results final
            raise ValueError("SlabTransfer(halo=...) needs own_box, the box of the slab's OWN points "
                             "(the one exchange_ghosts was called with)")
        if own_box is not None:
            box = own_box.to(torch.float64).to(dev)
        else:
            box = torch.stack([engine.bbox_lo, engine.bbox_hi]).to(torch.float64)
        self.own_box = box.clone()
        if engine.n_points == 0:           # an empty slab can never hold a neighbour
            box[0].fill_(float("inf"))
            box[1].fill_(float("-inf"))
        if self.world > 1:
            boxes = [torch.empty_like(box) for _ in range(self.world)]
            dist.all_gather(boxes, box, group=group)
            self.boxes = torch.stack(boxes)          # [R, 2, 3]
        else:
            self.boxes = box[None]
        self.device = dev
        self.stats = {}
        self.cap = 2048                       # halo rows per peer; doubled after an overflow
        self.boxes6 = self.boxes.reshape(self.world, 6).contiguous()

    def validate(self):
        """Checks the deferred overflow flags of the steps run with ``validate=False``.
        Returns True if every such step fitted its halo capacity (its results are exact);
        otherwise doubles the capacity and returns False (the caller must redo those steps)."""
        if self.world == 1:
            return True
        acc = getattr(self, "_flag_acc", None)
        if acc is None:
            return True
        dist.all_reduce(acc, op=dist.ReduceOp.MAX, group=self.group)
        bad = int(acc.item())
        acc.zero_()
        if bad:
            self.cap *= 2
        return not bad

    def _transfer_ghost(self, q, k, radius, want_d2, validate=True, r2pq=None):
        """Owner-only step on the ghost-augmented index + the check that no sample's k-th
        neighbour ball leaves the ghost zone.  Returns None when the step must be redone with
        the exchange (only possible with validate=True; otherwise validate() reports it)."""
        eng = self.engine
        kw = {} if r2pq is None else {"radius2_per_query": r2pq}
        try:
            _, out = eng.query(q, k, radius=radius, outputs=True, want_d2=True, want_cand=False, **kw)
        except TypeError:            # engines without the want_cand switch (test stand-ins)
            _, out = eng.query(q, k, radius=radius, outputs=True, want_d2=True, **kw)
        if q.shape[0] and hasattr(eng, "ghost_check"):       # one small kernel on the CUDA engine
            if not validate:
                # deferred check: the kernel ORs (atomically, so chunks on several streams may
                # share it) straight into the accumulated flag that validate() reduces
                if getattr(self, "_flag_acc", None) is None:
                    self._flag_acc = torch.zeros((1,), dtype=torch.int32, device=self.device)
                eng.ghost_check(q, out["d2"], k, radius, self.boxes6, self.rank, self.halo,
                                self._flag_acc)
                if not want_d2:
                    out.pop("d2")
                return out
            viol = torch.zeros((1,), dtype=torch.int32, device=self.device)
            eng.ghost_check(q, out["d2"], k, radius, self.boxes6, self.rank, self.halo, viol)
        elif q.shape[0]:
            kth = out["d2"][:, k - 1]
            if radius is not None and radius >= 0:
                kth = torch.clamp(kth, max=float(radius) * float(radius))
            # a sample must go to the exchange only if its ball reaches another slab's box AND
            # may stick out of the ghost zone around this slab
            e = torch.clamp(torch.maximum(self.own_box[0] - q, q - self.own_box[1]), min=0.0)
            leaves_zone = (torch.sqrt(kth) + torch.sqrt((e * e).sum(dim=1))) * (1.0 + 1e-9) > self.halo
            ep = torch.clamp(torch.maximum(self.boxes[None, :, 0] - q[:, None, :],
                                           q[:, None, :] - self.boxes[None, :, 1]), min=0.0)
            lbp = (ep * ep).sum(dim=2) * (1.0 - 1e-12)              # [m, R]
            lbp[:, self.rank] = float("inf")
            reaches_peer = (lbp <= kth[:, None]).any(dim=1)
            viol = (leaves_zone & reaches_peer).any().to(torch.int32).view(1)
        else:
            viol = torch.zeros((1,), dtype=torch.int32, device=self.device)
        if not want_d2:
            out.pop("d2")
        if not validate:
            if getattr(self, "_flag_acc", None) is None:
                self._flag_acc = torch.zeros((1,), dtype=torch.int32, device=self.device)
            torch.maximum(self._flag_acc, viol, out=self._flag_acc)
            return out
        dist.all_reduce(viol, op=dist.ReduceOp.MAX, group=self.group)
        return None if int(viol.item()) else out

    def _transfer_fast(self, q, k, radius, want_d2, validate=True, r2pq=None):
        """Fixed-capacity exchange on the CUDA engine: route kernel -> all_to_all -> bounded halo
        search -> all_to_all -> per-peer merge kernels; one deferred overflow check at the end."""
        eng, R, cap = self.engine, self.world, self.cap
        own, out = eng.query(q, k, radius=radius, outputs=True, want_d2=want_d2,
                             **({} if r2pq is None else {"radius2_per_query": r2pq}))
        h = eng.halo_buffers(R, cap, k)
        h["flag"].zero_()
        eng.halo_route(q, own, k, radius, self.boxes6, self.rank, cap, h)
        dist.all_to_all_single(h["recv"].view(R, -1), h["send"].view(R, -1), group=self.group)
        eng.halo_prepare(h, R, cap)
        eng.tree.query(h["hq"], k, radius2_per_query=h["hr2"], cand=h["hcand"].view(-1))
        dist.all_to_all_single(h["back"].view(R, -1), h["hcand"].view(R, -1), group=self.group)
        for r in range(R):
            if r != self.rank:
                eng.halo_merge(own, h, r, cap, k, out)
        if not validate:                      # deferred: accumulate, checked by validate()
            if getattr(self, "_flag_acc", None) is None:
                self._flag_acc = torch.zeros_like(h["flag"])
            torch.maximum(self._flag_acc, h["flag"], out=self._flag_acc)
            self._last_counts = h["counts"]
            return out
        # every rank must take the same path: agree on the overflow flag (one tiny all-reduce,
        # the only host synchronisation of the step)
        dist.all_reduce(h["flag"], op=dist.ReduceOp.MAX, group=self.group)
        if int(h["flag"].item()):             # some peer block did not fit: exact path, larger cap
            self.cap *= 2
            return None
        self._last_counts = h["counts"]
        return out

    def crossing_count(self):
        """Samples routed to other slabs in the last fast-path step (synchronises)."""
        c = getattr(self, "_last_counts", None)
        return None if c is None else int(c.sum().item())

    def _all_to_all(self, send, send_counts, width, dtype):
        """Variable-size exchange of rows ([n, width] tensors ordered by destination rank)."""
        dev = self.device
        sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=self.group)
        recv_counts = rc.tolist()
        recv = torch.empty((sum(recv_counts), width), dtype=dtype, device=dev)
        dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=recv_counts,
                               input_split_sizes=list(send_counts), group=self.group)
        return recv, recv_counts

    def transfer_host(self, q_host, k, out, radius=None, chunks=3):
        """Host-buffer entry point of the sharded path: ``q_host`` float64 [m,3] (pinned for true
        overlap) holds the samples this rank owns, ``out`` maps "idx" [m,k] int32, "rgba" [m,4]
        uint8, "normal" [m,3] float32 to (pinned) host tensors that receive the results.
        On the CUDA engine this is one call of the C ABI (``pt_transfer_slab``: the H2D copy of
        the samples, the owner step with the ghost-zone check and the D2H copy of the results are
        pipelined chunk by chunk inside the library) followed by ONE tiny all-reduce in which the
        ranks agree that nobody needs the exchange; otherwise (or on other engines, through
        ``chunks`` torch-stream pieces) the batch is redone through the exchange path.
        Returns after everything has landed."""
        m = q_host.shape[0]
        dev = self.device
        if self.world == 1 or self.halo is None or not getattr(self.engine, "fast", False) or m == 0:
            r = self.transfer(q_host.to(dev, non_blocking=True), k, radius=radius)
            for name in ("idx", "rgba", "normal"):
                out[name].copy_(r[name], non_blocking=True)
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            return
        tree = getattr(self.engine, "tree", None)
        if tree is not None and hasattr(tree, "transfer_slab") and q_host.dtype == torch.float64 \
                and q_host.is_contiguous() and all(out[n].is_contiguous() for n in ("idx", "rgba", "normal")):
            # the C host path: H2D / owner step + ghost check / D2H pipelined chunk by chunk
            # inside the library (pt_transfer_slab), then ONE agreement across the ranks
            if getattr(self, "_boxes6_host", None) is None:
                self._boxes6_host = self.boxes6.cpu().numpy()
            ok = tree.transfer_slab(q_host.data_ptr(), True, m, k, radius, self._boxes6_host,
                                    self.rank, self.halo, out["idx"].data_ptr(),
                                    out["rgba"].data_ptr(), out["normal"].data_ptr())
            bad = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
            if int(bad.item()) == 0:
                self.stats = {"crossing": 0, "sent": 0, "received": 0, "path": "ghost-zone (host)",
                              "halo": self.halo}
                return
        else:
            chunks = max(1, min(int(chunks), m))
            step = -(-m // chunks)
            if getattr(self, "_host_streams", None) is None:
                self._host_streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
            cur = torch.cuda.current_stream(dev)
            keep = []
            for c in range(chunks):
                lo, hi = c * step, min(m, (c + 1) * step)
                st = self._host_streams[c % len(self._host_streams)]
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    qd = q_host[lo:hi].to(dev, non_blocking=True)
                    r = self.transfer(qd, k, radius=radius, validate=False)
                    for name in ("idx", "rgba", "normal"):
                        out[name][lo:hi].copy_(r[name], non_blocking=True)
                    keep.append((qd, r))
            for st in self._host_streams:
                cur.wait_stream(st)
            torch.cuda.synchronize(dev)
            if self.validate():
                return
        # some sample needed the exchange: exact path, whole batch
        r = self.transfer(q_host.to(dev, non_blocking=True), k, radius=radius)
        for name in ("idx", "rgba", "normal"):
            out[name].copy_(r[name], non_blocking=True)
        torch.cuda.synchronize(dev)

    def transfer(self, q, k, radius=None, want_d2=False, validate=True, r2pq=None):
        """q float64 [m,3] on the engine's device (samples owned by this rank).
        Returns dict(idx int32 [m,k] global ids, rgba uint8 [m,4], normal float32 [m,3][, d2]).
        ``validate=False`` (CUDA engine only) skips the per-step overflow agreement, making the
        step fully asynchronous; the caller must then call ``validate()`` before trusting it."""
        eng, dev, R = self.engine, self.device, self.world
        m = q.shape[0]
        if R > 1 and self.halo is not None:
            out = self._transfer_ghost(q, k, radius, want_d2, validate, r2pq)
            if out is not None:
                self.stats = {"crossing": 0, "sent": 0, "received": 0, "path": "ghost-zone",
                              "halo": self.halo}
                return out
        if R > 1 and getattr(eng, "fast", False):
            out = self._transfer_fast(q, k, radius, want_d2, validate, r2pq)
            if out is not None:
                self.stats = {"crossing": None, "sent": None, "received": None, "path": "fast",
                              "cap": self.cap}
                return out
        r2 = float("inf") if (radius is None or radius < 0) else float(radius) * float(radius)
        own, out = eng.query(q, k, radius=radius, outputs=True, want_d2=want_d2,
                             **({} if r2pq is None else {"radius2_per_query": r2pq}))   # [m, k, 32]
        if R == 1:
            self.stats = {"crossing": 0, "sent": 0, "received": 0}
            return out
        # bound = squared distance of the k-th local neighbour (inf while the list is short)
        bound = torch.clamp(cand_d2(own)[:, k - 1], max=r2) if m else q.new_empty((0,))
        send_rows, send_counts, send_sel = [], [], []
        for r in range(R):
            if r == self.rank or m == 0:
                sel = torch.empty((0,), dtype=torch.int64, device=dev)
            else:
                lb = box_lower_bound2(q, self.boxes[r, 0], self.boxes[r, 1])
                sel = torch.nonzero(lb <= bound).view(-1)
            send_sel.append(sel)
            send_counts.append(int(sel.numel()))
            send_rows.append(torch.cat([q[sel], bound[sel, None]], dim=1))
        send = torch.cat(send_rows) if send_rows else q.new_empty((0, 4))
        recv, recv_counts = self._all_to_all(send, send_counts, 4, torch.float64)
        # halo search for the other ranks' samples, bounded by their k-th distance
        rq = recv[:, :3].contiguous()
        rb = recv[:, 3].contiguous()
        halo, _ = eng.query(rq, k, radius=None, radius2_per_query=rb)  # [n_recv, k, 32]
        back, back_counts = self._all_to_all(halo.view(-1, k * CAND_BYTES), recv_counts,
                                             k * CAND_BYTES, torch.uint8)
        assert back_counts == send_counts
        # merge: own list + one list per rank that was asked
        cross = torch.unique(torch.cat(send_sel)) if sum(send_counts) else \
            torch.empty((0,), dtype=torch.int64, device=dev)
        nc = int(cross.numel())
        if nc:
            lists = empty_cand((R, nc, k), dev)
            pos = torch.full((m,), -1, dtype=torch.int64, device=dev)
            pos[cross] = torch.arange(nc, device=dev)
            lists[self.rank] = own[cross]
            off = 0
            for r in range(R):
                c = send_counts[r]
                if c:
                    lists[r, pos[send_sel[r]]] = back[off:off + c].view(c, k, CAND_BYTES)
                off += c
            merged = eng.merge(lists, k, want_d2=want_d2)
            for name, t in merged.items():
                out[name][cross] = t
        self.stats = {"crossing": nc, "sent": int(sum(send_counts)),
                      "received": int(sum(recv_counts)), "path": "exact-size"}
        return out
