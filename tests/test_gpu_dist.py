"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on a single-GPU box): slab-sharded
transfer over NCCL must be bit-identical to the oracle on the whole cloud."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, k, radius):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import __graft_entry__ as ge
        from oracle import pto
        pkg = ge.package()
        P = pkg.synth.cloud_host(150_000, seed=17, side=60.0)
        V = pkg.synth.samples_host(60, side=60.0)
        order = np.argsort(P["ver"][:, 0], kind="stable")
        cuts = np.linspace(0, len(P), world + 1).astype(int)
        mine = np.sort(order[cuts[rank]:cuts[rank + 1]]).astype(np.int32)
        x_cut = [-np.inf] + [P["ver"][order[c], 0] for c in cuts[1:-1]] + [np.inf]
        own_q = np.nonzero((V["ver"][:, 0] >= x_cut[rank]) & (V["ver"][:, 0] < x_cut[rank + 1]))[0]
        sub = P[mine]
        pos = torch.zeros((len(mine), 4), dtype=torch.float32, device=dev)
        pos[:, :3] = torch.from_numpy(sub["ver"].astype(np.float32)).to(dev)
        attrs = np.zeros(len(mine), dtype=pkg.ATTR_DTYPE)
        attrs["nx"], attrs["ny"], attrs["nz"] = sub["normal"][:, 0], sub["normal"][:, 1], sub["normal"][:, 2]
        attrs["rgba"][:, :3] = sub["color"]
        attrs["rgba"][:, 3] = 255
        tree = pkg.DeviceTree(pos, torch.from_numpy(attrs.view(np.uint8).reshape(-1, 16)).to(dev),
                              torch.from_numpy(mine).to(dev))
        st = pkg.dist.SlabTransfer(pkg.dist.CudaSlabEngine(tree))
        q = torch.from_numpy(np.ascontiguousarray(V["ver"][own_q])).to(dev)
        out = st.transfer(q, k, radius=radius, want_d2=True)
        torch.cuda.synchronize()
        ref_idx, ref_d2 = pto.KdTree(P).knn(V[own_q], k, radius=-1.0 if radius is None else radius)
        ref_rgba, ref_nrm = pto.blend(P, ref_idx, ref_d2)
        assert np.array_equal(out["idx"].cpu().numpy(), ref_idx), f"rank {rank} idx"
        assert np.array_equal(out["d2"].cpu().numpy(), ref_d2), f"rank {rank} d2"
        assert np.array_equal(out["rgba"].cpu().numpy(), ref_rgba), f"rank {rank} rgba"
        assert np.allclose(out["normal"].cpu().numpy(), ref_nrm, rtol=1e-5, atol=1e-7)
        assert st.stats["path"] == "fast"             # fixed-capacity exchange, no overflow
        assert st.crossing_count() < len(own_q)       # only a sliver is exchanged
        # the exact-size path (taken after a capacity overflow) must agree bit for bit
        st.engine.fast = False
        out2 = st.transfer(q, k, radius=radius, want_d2=True)
        torch.cuda.synchronize()
        assert st.stats["path"] == "exact-size"
        for name in ("idx", "d2", "rgba", "normal"):
            assert torch.equal(out[name], out2[name]), name
        tree.close()
        # ---- slab + ghost zone: points near the other slabs exchanged once at build time, then
        # owner-only steps with no collective; still bit-identical to the whole-cloud oracle
        halo = 3.0 if radius is None else radius + 1.0
        own_box = pkg.dist.points_box(pos)
        boxes = pkg.dist.gather_boxes(own_box)
        gpos, gattr, gids = pkg.dist.exchange_ghosts(
            pos, torch.from_numpy(attrs.view(np.uint8).reshape(-1, 16)).to(dev),
            torch.from_numpy(mine).to(dev), boxes, halo)
        assert gpos.shape[0] > pos.shape[0]
        gtree = pkg.DeviceTree(gpos, gattr, gids)
        gst = pkg.dist.SlabTransfer(pkg.dist.CudaSlabEngine(gtree), own_box=own_box, halo=halo)
        gout = gst.transfer(q, k, radius=radius, want_d2=True)
        torch.cuda.synchronize()
        assert gst.stats["path"] == "ghost-zone", gst.stats
        for name in ("idx", "d2", "rgba", "normal"):
            assert torch.equal(out[name], gout[name]), name
        # host-buffer entry point: chunks pipelined over streams, one deferred validation
        q_pin = torch.from_numpy(np.ascontiguousarray(V["ver"][own_q])).pin_memory()
        host = {"idx": torch.empty((len(own_q), k), dtype=torch.int32).pin_memory(),
                "rgba": torch.empty((len(own_q), 4), dtype=torch.uint8).pin_memory(),
                "normal": torch.empty((len(own_q), 3), dtype=torch.float32).pin_memory()}
        gst.transfer_host(q_pin, k, host, radius=radius, chunks=3)
        for name in ("idx", "rgba", "normal"):
            assert torch.equal(host[name], out[name].cpu()), "transfer_host " + name
        # ... and when the deferred validation fails the batch is redone through the exchange
        gst_n = pkg.dist.SlabTransfer(pkg.dist.CudaSlabEngine(gtree), own_box=own_box, halo=0.01)
        for t in host.values():
            t.zero_()
        gst_n.transfer_host(q_pin, k, host, radius=radius, chunks=2)
        for name in ("idx", "rgba", "normal"):
            assert torch.equal(host[name], out[name].cpu()), "transfer_host fallback " + name
        # a ghost zone that is too narrow falls back to the exchange (duplicates dropped)
        gst2 = pkg.dist.SlabTransfer(pkg.dist.CudaSlabEngine(gtree), own_box=own_box, halo=0.01)
        gout2 = gst2.transfer(q, k, radius=radius, want_d2=True)
        torch.cuda.synchronize()
        assert gst2.stats["path"] == "fast", gst2.stats
        for name in ("idx", "d2", "rgba", "normal"):
            assert torch.equal(out[name], gout2[name]), name
        gtree.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,radius", [(16, None), (8, 0.6)])
def test_two_gpu_slab_transfer(k, radius):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 300, k, radius), nprocs=world, join=True)
