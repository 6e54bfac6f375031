"""Synthetic workloads of BASELINE.json's configs.

The reference ships no sample data, so clouds and mesh samples are generated on the device by
``pt_synth_*`` (csrc/pt_synth.cu, built into the bench-only libpt_synth_b200.so): a noisy heightfield scan over ``[0, L]^2`` or a skewed
cluster cloud, Philox4x32-10 keyed by ``seed`` with the global point index as counter.
"""
import ctypes
from dataclasses import dataclass

import numpy as np

from . import api

L_DOMAIN = 1000.0
SEED0 = 20261018


@dataclass(frozen=True)
class Workload:
    name: str
    n_points: int      # per slab (rank)
    gu: int            # samples along u per slab
    gv: int            # samples along v
    k: int
    kind: int = api.SYNTH_HEIGHTFIELD
    radius: float = None
    center: bool = False
    seed: int = SEED0
    sigma: float = 0.01

    @property
    def n_samples(self):
        return self.gu * self.gv


# BASELINE.json configs[0..4]; n_points / samples are totals for the single-slab case.
CONFIGS = {
    "cfg1": Workload("cfg1: 1M-point cloud -> 10k-vertex mesh, k=8", 1_000_000, 100, 100, 8,
                     seed=SEED0 + 1),
    "cfg2": Workload("cfg2: 50M-point cloud -> 200k-vertex mesh, k=16", 50_000_000, 448, 448, 16,
                     seed=SEED0 + 2),
    "cfg3": Workload("cfg3: 300M-point scan -> 1M-vertex mesh, k=16", 300_000_000, 1000, 1000, 16,
                     seed=SEED0 + 3),
    "cfg4": Workload("cfg4: 1B-point cloud -> 4M texel samples, k=32", 1_000_000_000, 2048, 2048,
                     32, center=True, seed=SEED0 + 4),
    "cfg5": Workload("cfg5: skewed 200M-point cloud -> 1M samples, k=16, R=0.25", 200_000_000,
                     1000, 1000, 16, kind=api.SYNTH_SKEWED, radius=0.25, seed=SEED0 + 5,
                     sigma=L_DOMAIN / 2000.0),
}


def cloud_device(n, seed, u0=0.0, u1=L_DOMAIN, v0=0.0, v1=L_DOMAIN, kind=api.SYNTH_HEIGHTFIELD,
                 sigma=0.01, first_index=0, device=None, want_attrs=True):
    """Returns (pos float32 [n,4], attrs uint8 [n,16]) CUDA tensors."""
    import torch
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    pos = torch.empty((n, 4), dtype=torch.float32, device=device)
    attrs = torch.empty((n, 16), dtype=torch.uint8, device=device) if want_attrs else None
    sp = api.SynthParams(kind=int(kind), seed=int(seed), first_index=int(first_index),
                         u0=u0, u1=u1, v0=v0, v1=v1, sigma=sigma)
    with torch.cuda.device(device):
        api._check(api.synth_lib().pt_synth_cloud_device(api._tptr(pos), api._tptr(attrs), n,
                                                   ctypes.byref(sp), api._stream_ptr()),
                   "pt_synth_cloud_device")
    return pos, attrs


def samples_device(gu, gv, u0=0.0, u1=L_DOMAIN, v0=0.0, v1=L_DOMAIN, center=False, device=None):
    """Returns float64 [gu*gv, 3] CUDA tensor of sample positions on the noise-free surface."""
    import torch
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    q = torch.empty((gu * gv, 3), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        api._check(api.synth_lib().pt_synth_samples_device(api._tptr(q), gu, gv, u0, u1, v0, v1,
                                                     1 if center else 0, api._stream_ptr()),
                   "pt_synth_samples_device")
    return q


def points_to_host(pos, attrs):
    """Device SoA -> host array of 80-byte ``Point`` records (for the host API / the oracle)."""
    import torch
    n = pos.shape[0]
    raw = torch.empty((n, 80), dtype=torch.uint8, device=pos.device)
    with torch.cuda.device(pos.device):
        api._check(api.synth_lib().pt_synth_pack_points_device(api._tptr(pos), api._tptr(attrs), n,
                                                         api._tptr(raw), api._stream_ptr()),
                   "pt_synth_pack_points_device")
    return raw.cpu().numpy().view(api.POINT_DTYPE).reshape(-1)


def queries_to_host(q, pinned=False):
    import torch
    m = q.shape[0]
    raw = torch.empty((m, 80), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        api._check(api.synth_lib().pt_synth_pack_queries_device(api._tptr(q), m, api._tptr(raw),
                                                          api._stream_ptr()),
                   "pt_synth_pack_queries_device")
    if pinned:
        host = torch.empty((m, 80), dtype=torch.uint8, pin_memory=True)
        host.copy_(raw)
        return host
    return raw.cpu().numpy().view(api.POINT_DTYPE).reshape(-1)


def grid_faces(gu, gv):
    """2(gu-1)(gv-1) triangles of the gu x gv vertex grid (row-major), int32 [F,3]."""
    j, i = np.meshgrid(np.arange(gv - 1), np.arange(gu - 1), indexing="ij")
    a = (j * gu + i).ravel()
    f = np.stack([np.stack([a, a + 1, a + gu], 1), np.stack([a + 1, a + gu + 1, a + gu], 1)], 1)
    return f.reshape(-1, 3).astype(np.int32)


# ---- host (numpy) generator for CPU-only tests; NOT bit-identical to the device generator ----
def surface_np(x, y):
    A = (12.0, 6.0, 2.5, 0.8); F = (0.021, 0.047, 0.11, 0.31); G = (0.017, 0.039, 0.13, 0.27)
    P = (0.3, 1.7, 2.9, 0.5); Q = (1.1, 0.2, 4.1, 3.3)
    z = np.zeros_like(x)
    for a, f, g, p, q in zip(A, F, G, P, Q):
        z = z + a * np.sin(f * x + p) * np.sin(g * y + q)
    return z


def cloud_host(n, seed, side=100.0, sigma=0.01):
    rng = np.random.default_rng(seed)
    u, v = rng.random(n) * side, rng.random(n) * side
    z = surface_np(u, v) + rng.normal(0.0, sigma, n)
    xyz = np.stack([u, v, z], 1).astype(np.float32).astype(np.float64)
    nrm = rng.standard_normal((n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    col = rng.integers(0, 256, (n, 3))
    return api.make_points(xyz, normal=nrm.astype(np.float32), color=col)


def samples_host(g, side=100.0):
    t = np.linspace(0.0, side, g)
    u, v = np.meshgrid(t, t)
    u, v = u.ravel(), v.ravel()
    xyz = np.stack([u, v, surface_np(u, v)], 1).astype(np.float32).astype(np.float64)
    return api.make_points(xyz)
