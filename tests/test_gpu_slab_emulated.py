"""Slab-sharded path on ONE GPU: R slab indexes live on the same device and the collectives of
the exchange are emulated by handing the buffers from "rank" to "rank" in Python, so that every
multi-GPU kernel and ABI entry -- pt_transfer_slab, pt_ghost_check_device, pt_halo_route_device,
pt_halo_prepare_device, pt_halo_merge_device, pt_route_samples_device, pt_scatter_rows_device --
is compared with the oracle on the WHOLE cloud where the driver's single-GPU box can see it
(the reference has no distributed path: its query loop is
src/pointsTransfer.cpp:465-479).  Both branches run: "the ghost zone suffices" and "the step
needs the exchange".
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(autouse=True)
def _default_options(pkg):
    yield
    pkg.set_option("knn_variant", -1)


def _attrs_of(pkg, sub):
    a = np.zeros(len(sub), dtype=pkg.ATTR_DTYPE)
    a["nx"], a["ny"], a["nz"] = sub["normal"][:, 0], sub["normal"][:, 1], sub["normal"][:, 2]
    a["rgba"][:, :3] = sub["color"]
    a["rgba"][:, 3] = 255
    return a


class Slabs:
    """R slabs of one host cloud cut along x at point-count quantiles, optionally with a ghost
    zone of width `halo` (the other slabs' points within halo of the slab's own box)."""

    def __init__(self, pkg, torch, P, R, halo=None):
        self.pkg, self.torch, self.R = pkg, torch, R
        x = P["ver"][:, 0]
        order = np.argsort(x, kind="stable")
        cuts = np.linspace(0, len(P), R + 1).astype(int)
        self.x_cut = np.array([-np.inf] + [x[order[c]] for c in cuts[1:-1]] + [np.inf])
        self.own_ids, self.boxes, self.trees, self.engines = [], [], [], []
        for r in range(R):
            ids = np.sort(order[cuts[r]:cuts[r + 1]]).astype(np.int32)
            self.own_ids.append(ids)
            v = P["ver"][ids]
            self.boxes.append(np.concatenate([v.min(0), v.max(0)]))
        self.boxes6 = torch.from_numpy(np.stack(self.boxes)).cuda()
        for r in range(R):
            ids = self.own_ids[r]
            if halo is not None:
                lo, hi = self.boxes[r][:3], self.boxes[r][3:]
                e = np.maximum(np.maximum(lo - P["ver"], P["ver"] - hi), 0.0)
                near = np.nonzero((e * e).sum(1) <= halo * halo * (1 + 1e-9))[0].astype(np.int32)
                ids = np.union1d(ids, near).astype(np.int32)        # ascending global ids
            sub = P[ids]
            pos = torch.zeros((len(ids), 4), dtype=torch.float32, device="cuda")
            pos[:, :3] = torch.from_numpy(sub["ver"].astype(np.float32)).cuda()
            t = pkg.DeviceTree(pos, torch.from_numpy(_attrs_of(pkg, sub).view(np.uint8).reshape(-1, 16)).cuda(),
                               torch.from_numpy(ids).cuda())
            self.trees.append(t)
            self.engines.append(pkg.dist.CudaSlabEngine(t))

    def owner_of(self, V):
        return np.searchsorted(self.x_cut, V["ver"][:, 0], side="right") - 1

    def close(self):
        for t in self.trees:
            t.close()


def _exchange_step(S, queries, k, radius, cap):
    """The fixed-capacity exchange of dist.SlabTransfer._transfer_fast with the two all_to_all
    calls replaced by copies between the emulated ranks.  queries[r]: float64 [m_r, 3] CUDA."""
    torch, R = S.torch, S.R
    own, out, bufs = [], [], []
    for r in range(R):
        c, o = S.engines[r].query(queries[r], k, radius=radius, outputs=True, want_d2=True)
        h = S.engines[r].halo_buffers(R, cap, k)
        h["flag"].zero_()
        S.engines[r].halo_route(queries[r], c, k, radius, S.boxes6, r, cap, h)
        own.append(c); out.append(o); bufs.append(h)
    for d in range(R):                                   # all_to_all #1
        for r in range(R):
            bufs[d]["recv"][r].copy_(bufs[r]["send"][d])
    for d in range(R):                                   # bounded halo search on the receiver
        S.engines[d].halo_prepare(bufs[d], R, cap)
        S.trees[d].query(bufs[d]["hq"], k, radius2_per_query=bufs[d]["hr2"], cand=bufs[d]["hcand"].view(-1))
    for r in range(R):                                   # all_to_all #2
        for d in range(R):
            bufs[r]["back"][d].copy_(bufs[d]["hcand"].view(R, cap, k, 32)[r])
    crossing = 0
    for r in range(R):
        for d in range(R):
            if d != r:
                S.engines[r].halo_merge(own[r], bufs[r], d, cap, k, out[r])
        assert int(bufs[r]["flag"].item()) == 0, "halo capacity overflow"
        crossing += int(bufs[r]["counts"].sum().item())
    torch.cuda.synchronize()
    return out, crossing


def _check(out, ref, sel):
    ref_idx, ref_d2, ref_rgba, ref_nrm = ref
    assert np.array_equal(out["idx"].cpu().numpy(), ref_idx[sel])
    assert np.array_equal(out["d2"].cpu().numpy(), ref_d2[sel])
    assert np.array_equal(out["rgba"].cpu().numpy(), ref_rgba[sel])
    assert np.allclose(out["normal"].cpu().numpy(), ref_nrm[sel], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("variant", [-1, 2])
@pytest.mark.parametrize("R,k,radius", [(2, 16, None), (3, 8, 0.6), (4, 32, None)])
def test_exchange_path_emulated(R, k, radius, variant, pkg, pto, torch_cuda):
    """No ghost zone: every sample whose k-th-neighbour ball reaches another slab's box goes
    through halo route -> bounded halo search -> halo merge."""
    torch = torch_cuda
    pkg.set_option("knn_variant", variant)
    P = pkg.synth.cloud_host(150_000, seed=17, side=60.0)
    V = pkg.synth.samples_host(60, side=60.0)
    ref_idx, ref_d2 = pto.KdTree(P).knn(V, k, radius=-1.0 if radius is None else radius)
    ref = (ref_idx, ref_d2) + pto.blend(P, ref_idx, ref_d2)
    S = Slabs(pkg, torch, P, R)
    owner = S.owner_of(V)
    sels = [np.nonzero(owner == r)[0] for r in range(R)]
    queries = [torch.from_numpy(np.ascontiguousarray(V["ver"][s])).cuda() for s in sels]
    out, crossing = _exchange_step(S, queries, k, radius, cap=2048)
    assert 0 < crossing < len(V)            # a sliver of boundary samples really was exchanged
    for r in range(R):
        _check(out[r], ref, sels[r])
    S.close()


@pytest.mark.parametrize("R,k,radius", [(2, 16, None), (4, 8, 0.6)])
def test_ghost_zone_path_emulated(R, k, radius, pkg, pto, torch_cuda):
    """Ghost-augmented slab indexes through the host-buffer entry pt_transfer_slab: a wide halo
    makes every step final with no exchange; a narrow one must raise needs_exchange, and the
    exchange (with id de-duplication against the ghost copies) then restores exactness."""
    torch = torch_cuda
    P = pkg.synth.cloud_host(150_000, seed=23, side=60.0)
    V = pkg.synth.samples_host(60, side=60.0)
    ref_idx, ref_d2 = pto.KdTree(P).knn(V, k, radius=-1.0 if radius is None else radius)
    ref = (ref_idx, ref_d2) + pto.blend(P, ref_idx, ref_d2)
    for halo, expect_final in ((3.0 if radius is None else radius + 1.0, True), (0.01, False)):
        S = Slabs(pkg, torch, P, R, halo=halo)
        owner = S.owner_of(V)
        boxes_host = np.stack(S.boxes)
        any_needs = False
        for r in range(R):
            sel = np.nonzero(owner == r)[0]
            qh = np.ascontiguousarray(V["ver"][sel])
            m = len(sel)
            o = {"idx": np.empty((m, k), np.int32), "d2": np.empty((m, k), np.float64),
                 "rgba": np.empty((m, 4), np.uint8), "normal": np.empty((m, 3), np.float32)}
            final = S.trees[r].transfer_slab(qh.ctypes.data, True, m, k, radius, boxes_host, r, halo,
                                             o["idx"].ctypes.data, o["rgba"].ctypes.data,
                                             o["normal"].ctypes.data, d2_ptr=o["d2"].ctypes.data)
            any_needs |= not final
            if final:
                _check({n: torch.from_numpy(a) for n, a in o.items()}, ref, sel)
            # the device-side check agrees with the host entry's verdict
            flag = torch.zeros(1, dtype=torch.int32, device="cuda")
            S.engines[r].ghost_check(torch.from_numpy(qh).cuda(), torch.from_numpy(o["d2"]).cuda(), k, radius,
                                     S.boxes6, r, halo, flag)
            assert bool(flag.item()) == (not final)
        assert any_needs == (not expect_final)
        if not expect_final:
            sels = [np.nonzero(owner == r)[0] for r in range(R)]
            queries = [torch.from_numpy(np.ascontiguousarray(V["ver"][s])).cuda() for s in sels]
            out, crossing = _exchange_step(S, queries, k, radius, cap=2048)
            assert crossing > 0
            for r in range(R):
                _check(out[r], ref, sels[r])
        S.close()


def test_route_and_scatter_kernels(pkg, torch_cuda):
    """pt_route_samples_device / pt_scatter_rows_device: every sample lands in the block of the
    slab whose x-range holds it, unused rows stay NaN / -1, and scattering the rows back by
    `sel` is the identity."""
    torch = torch_cuda
    rng = np.random.default_rng(3)
    m, R, cap = 10_000, 4, 4096
    q = torch.from_numpy(rng.random((m, 3)) * 100.0).cuda()
    cuts = torch.tensor([-np.inf, 20.0, 55.0, 80.0, np.inf], dtype=torch.float64, device="cuda")
    send, sel, counts, overflow = pkg.dist.route_samples(q, cuts, cap)
    torch.cuda.synchronize()
    assert int(overflow.item()) == 0
    owner = np.searchsorted(cuts.cpu().numpy(), q[:, 0].cpu().numpy(), side="right") - 1
    assert counts.cpu().numpy().tolist() == np.bincount(owner, minlength=R).tolist()
    send_h, sel_h = send.cpu().numpy(), sel.cpu().numpy()
    for r in range(R):
        c = int(counts[r])
        assert np.all(sel_h[r, c:] == -1) and np.all(np.isnan(send_h[r, c:, 3]))
        assert set(sel_h[r, :c].tolist()) == set(np.nonzero(owner == r)[0].tolist())
        assert np.array_equal(send_h[r, :c, :3], q.cpu().numpy()[sel_h[r, :c]])
        assert np.all(np.isinf(send_h[r, :c, 3]))
    back = torch.zeros((m, 3), dtype=torch.float64, device="cuda")
    pkg.dist.scatter_rows(send[:, :, :3].contiguous().view(R * cap, 3), sel.view(-1), back)
    torch.cuda.synchronize()
    assert torch.equal(back, q)
    # capacity overflow is reported, never silently dropped
    _, _, _, ovf = pkg.dist.route_samples(q, cuts, 16)
    assert int(ovf.item()) == 1
