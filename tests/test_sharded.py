"""world_size-2 / 3 `gloo` tests of the sharded public entry (``sharded.ShardedTransfer``):
samples arrive UNSORTED on every rank, are routed to the slab that owns them, answered there
(ghost zone or exchange + merge) and scattered back -- results must equal the oracle on the whole
cloud, in the caller's sample order.  CPU tensors + the oracle-backed engine of test_dist.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, halo, k, radius):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as ge
        from oracle import pto
        from test_dist import OracleEngine
        pkg = ge.package()
        P = pkg.synth.cloud_host(30_000, seed=5, side=40.0)
        V = pkg.synth.samples_host(36, side=40.0)
        order = np.argsort(P["ver"][:, 0], kind="stable")
        cuts_i = np.linspace(0, len(P), world + 1).astype(int)
        mine = np.sort(order[cuts_i[rank]:cuts_i[rank + 1]])
        x_cut = [-np.inf] + [P["ver"][order[c], 0] for c in cuts_i[1:-1]] + [np.inf]
        # this rank's part of the samples: a fixed pseudo-random share, in scrambled order
        perm = np.random.default_rng(9).permutation(len(V))
        share = perm[rank::world]
        own_box = None
        if halo is None:
            eng = OracleEngine(pkg, pto, P[mine], mine)
        else:
            lo, hi = P["ver"][mine].min(0), P["ver"][mine].max(0)
            e = np.maximum(np.maximum(lo - P["ver"], P["ver"] - hi), 0.0)
            ids = np.union1d(mine, np.nonzero((e * e).sum(1) <= halo * halo)[0])
            eng = OracleEngine(pkg, pto, P[ids], ids)
            own_box = torch.from_numpy(np.stack([lo, hi]))
        st = pkg.sharded.ShardedTransfer(eng, x_cut, k, len(share), own_box=own_box, halo=halo)
        q = torch.from_numpy(np.ascontiguousarray(V["ver"][share]))
        out = {"idx": torch.full((len(share), k), -7, dtype=torch.int32),
               "rgba": torch.zeros((len(share), 4), dtype=torch.uint8),
               "normal": torch.zeros((len(share), 3), dtype=torch.float32)}
        st.transfer(q, out, radius=radius)
        assert st.validate()
        ref_idx, ref_d2 = pto.knn_bruteforce(P, V[share], k, radius=-1.0 if radius is None else radius)
        ref_rgba, ref_nrm = pto.blend(P, ref_idx, ref_d2)
        assert np.array_equal(out["idx"].numpy(), ref_idx), f"rank {rank}: idx"
        assert np.array_equal(out["rgba"].numpy(), ref_rgba), f"rank {rank}: rgba"
        assert np.allclose(out["normal"].numpy(), ref_nrm, rtol=1e-5, atol=1e-7)
        assert st.stats()["routed_to_other_slabs"] > 0
        # host-buffer form, in two and three pieces: same results
        for pieces in (2, 3):
            host = {"idx": torch.full((len(share), k), -7, dtype=torch.int32),
                    "rgba": torch.zeros((len(share), 4), dtype=torch.uint8),
                    "normal": torch.zeros((len(share), 3), dtype=torch.float32)}
            st.transfer_host(q, host, radius=radius, pieces=pieces)
            assert st.validate()
            for name in ("idx", "rgba", "normal"):
                assert torch.equal(host[name], out[name]), (pieces, name)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,halo,k,radius", [(2, None, 8, None), (3, None, 16, 0.8), (2, 3.0, 8, None)])
def test_sharded_transfer_unsorted_samples(world, halo, k, radius):
    mp.spawn(_worker, args=(world, 29900 + os.getpid() % 90, halo, k, radius), nprocs=world, join=True)
