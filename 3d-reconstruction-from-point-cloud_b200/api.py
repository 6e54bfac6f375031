"""Host-side mirror of the reference's interface for the detail-transfer hot path.

The reference (horizon-research/3D-Reconstruction-From-Point-Cloud) exposes this path as
C++ objects in ``src/pointsTransfer.cpp``:

* ``Tree tree(points.begin(), points.end());``            (:259, typedefs :37-40)
* ``K_neighbor_search search(tree, query, K);``           (:474)
* iteration over ``(Point, squared distance)`` results    (:475-478)

with ``Point`` (src/Point.h) as the record type and ``Distance`` (src/Distance.h) as the
metric.  This module keeps those names and argument meanings and forwards everything to the
C ABI in ``include/points_transfer.h`` (``libpoints_transfer_b200.so``, hand-written sm_100a
CUDA).  There is no CPU fallback: if the library is missing or no CUDA device is present the
calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PT_LIB_NAME selects a diagnosis build of the same library (e.g. the -DPT_STATS one of `make stats`)
LIB_PATH = os.path.join(_HERE, os.environ.get("PT_LIB_NAME", "libpoints_transfer_b200.so"))

# Byte-for-byte mirror of the reference ``struct Point`` (src/Point.h:1-6), 80 bytes.
POINT_DTYPE = np.dtype(
    [("ver", "<f8", 3), ("normal", "<f8", 3), ("color", "<i4", 3), ("pad_", "<i4"),
     ("U", "<f8"), ("V", "<f8")]
)
assert POINT_DTYPE.itemsize == 80
ATTR_DTYPE = np.dtype([("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"), ("rgba", "u1", 4)])
CAND_DTYPE = np.dtype([("d2", "<f8"), ("id", "<i4"), ("rgba", "u1", 4), ("nx", "<f4"),
                       ("ny", "<f4"), ("nz", "<f4"), ("pad_", "<i4")])
assert ATTR_DTYPE.itemsize == 16 and CAND_DTYPE.itemsize == 32

PT_OK = 0
PT_ERR_NO_DEVICE = 3
PT_ERR_UNSUPPORTED = 5
PT_ERR_NOT_REPRESENTABLE = 6
PT_ERR_NON_FINITE = 7
PT_MAX_K = 32
COORD_AUTO, COORD_F32, COORD_F64 = 0, 1, 2
SYNTH_HEIGHTFIELD, SYNTH_SKEWED = 0, 1

# Every symbol include/points_transfer.h declares (the drop-in library) ...
ABI_SYMBOLS = (
    "pt_version", "pt_status_string", "pt_device_count", "pt_index_build", "pt_index_free",
    "pt_index_get_info", "pt_index_fallback_counts", "pt_knn", "pt_transfer", "pt_texture_render", "pt_transfer_slab", "pt_index_build_device", "pt_query_device",
    "pt_merge_device", "pt_halo_route_device", "pt_halo_prepare_device",
    "pt_halo_merge_device", "pt_ghost_check_device", "pt_route_samples_device",
    "pt_scatter_rows_device", "pt_sharded_build", "pt_sharded_free", "pt_sharded_get_info", "pt_sharded_knn",
    "pt_sharded_transfer", "pt_texture_render_lists", "pt_set_option", "pt_get_option", "pt_debug_stats", "pt_kernel_launch_count",
)
# ... and include/pt_synth.h (bench / test scaffolding, its own library)
SYNTH_SYMBOLS = ("pt_synth_cloud_device", "pt_synth_samples_device", "pt_synth_pack_points_device",
                 "pt_synth_pack_queries_device")
SYNTH_LIB_PATH = os.path.join(_HERE, "libpt_synth_b200.so")


class PointsTransferError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        super().__init__(f"{where}: status {status} ({status_string(status)})")


class BuildOpts(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("coord_mode", ctypes.c_int),
                ("ids", ctypes.c_void_p), ("reserved", ctypes.c_int * 8)]


class IndexInfo(ctypes.Structure):
    _fields_ = [("n_points", ctypes.c_uint64), ("n_leaves", ctypes.c_uint64),
                ("n_levels", ctypes.c_int), ("coord_mode", ctypes.c_int),
                ("device", ctypes.c_int), ("last_fallback_samples", ctypes.c_int),
                ("bbox_lo", ctypes.c_double * 3), ("bbox_hi", ctypes.c_double * 3),
                ("device_bytes", ctypes.c_uint64), ("build_ms", ctypes.c_float),
                ("last_query_ms", ctypes.c_float), ("last_h2d_ms", ctypes.c_float),
                ("last_d2h_ms", ctypes.c_float)]


class TextureStats(ctypes.Structure):
    _fields_ = [("triangles", ctypes.c_uint64), ("inside_points", ctypes.c_uint64),
                ("knn_ms", ctypes.c_float), ("draw_ms", ctypes.c_float), ("pad_ms", ctypes.c_float)]


class ShardedOpts(ctypes.Structure):
    _fields_ = [("n_devices", ctypes.c_int), ("devices", ctypes.POINTER(ctypes.c_int)), ("halo", ctypes.c_double),
                ("k_hint", ctypes.c_int), ("coord_mode", ctypes.c_int), ("reserved", ctypes.c_int * 8)]


class ShardedInfo(ctypes.Structure):
    _fields_ = [("n_slabs", ctypes.c_int), ("rebuilds", ctypes.c_int), ("halo", ctypes.c_double),
                ("n_points", ctypes.c_uint64), ("slab_points", ctypes.c_uint64 * 64),
                ("slab_ghosts", ctypes.c_uint64 * 64), ("slab_device", ctypes.c_int * 64)]


class SynthParams(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("seed", ctypes.c_uint64),
                ("first_index", ctypes.c_uint64), ("u0", ctypes.c_double),
                ("u1", ctypes.c_double), ("v0", ctypes.c_double), ("v1", ctypes.c_double),
                ("sigma", ctypes.c_double)]


_lib = None


def lib():
    """Load the CUDA library.  Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  points_transfer_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, i32, dbl = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double
    L.pt_version.restype = ctypes.c_char_p
    L.pt_status_string.restype = ctypes.c_char_p
    L.pt_status_string.argtypes = [i32]
    L.pt_device_count.restype = i32
    L.pt_index_build.restype = i32
    L.pt_index_build.argtypes = [vp, sz, ctypes.POINTER(BuildOpts), ctypes.POINTER(vp)]
    L.pt_index_free.restype = i32
    L.pt_index_free.argtypes = [vp]
    L.pt_index_get_info.restype = i32
    L.pt_index_get_info.argtypes = [vp, ctypes.POINTER(IndexInfo)]
    L.pt_index_fallback_counts.restype = i32
    L.pt_index_fallback_counts.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32 * 2)]
    L.pt_knn.restype = i32
    L.pt_knn.argtypes = [vp, vp, sz, i32, dbl, vp, vp]
    L.pt_transfer.restype = i32
    L.pt_transfer.argtypes = [vp, vp, sz, i32, dbl, vp, vp, vp, vp]
    L.pt_texture_render.restype = i32
    L.pt_texture_render.argtypes = [vp, vp, sz, vp, sz, i32, dbl, i32, i32, vp, ctypes.POINTER(TextureStats)]
    L.pt_transfer_slab.restype = i32
    L.pt_transfer_slab.argtypes = [vp, vp, i32, sz, i32, dbl, vp, i32, i32, dbl, vp, vp, vp, vp,
                                   ctypes.POINTER(i32)]
    L.pt_index_build_device.restype = i32
    L.pt_index_build_device.argtypes = [vp, i32, vp, vp, sz, i32, ctypes.POINTER(vp)]
    L.pt_query_device.restype = i32
    L.pt_query_device.argtypes = [vp, vp, sz, i32, dbl, vp, vp, vp, vp, vp, vp, vp]
    L.pt_merge_device.restype = i32
    L.pt_merge_device.argtypes = [vp, i32, sz, i32, vp, vp, vp, vp, vp, i32, vp]
    u32 = ctypes.c_uint32
    L.pt_halo_route_device.restype = i32
    L.pt_halo_route_device.argtypes = [vp, vp, sz, i32, dbl, vp, i32, i32, u32, vp, vp, vp, vp, vp]
    L.pt_halo_prepare_device.restype = i32
    L.pt_halo_prepare_device.argtypes = [vp, i32, u32, vp, vp, vp]
    L.pt_halo_merge_device.restype = i32
    L.pt_halo_merge_device.argtypes = [vp, vp, vp, vp, u32, i32, vp, vp, vp, vp, vp]
    L.pt_ghost_check_device.restype = i32
    L.pt_ghost_check_device.argtypes = [vp, vp, sz, i32, dbl, vp, i32, i32, dbl, vp, vp]
    L.pt_route_samples_device.restype = i32
    L.pt_route_samples_device.argtypes = [vp, sz, vp, i32, u32, vp, vp, vp, vp, vp]
    L.pt_scatter_rows_device.restype = i32
    L.pt_scatter_rows_device.argtypes = [vp, vp, sz, u32, vp, vp]
    L.pt_sharded_build.restype = i32
    L.pt_sharded_build.argtypes = [vp, sz, ctypes.POINTER(ShardedOpts), ctypes.POINTER(vp)]
    L.pt_sharded_free.restype = i32
    L.pt_sharded_free.argtypes = [vp]
    L.pt_sharded_get_info.restype = i32
    L.pt_sharded_get_info.argtypes = [vp, ctypes.POINTER(ShardedInfo)]
    L.pt_sharded_knn.restype = i32
    L.pt_sharded_knn.argtypes = [vp, vp, sz, i32, dbl, vp, vp]
    L.pt_sharded_transfer.restype = i32
    L.pt_sharded_transfer.argtypes = [vp, vp, sz, i32, dbl, vp, vp, vp, vp]
    L.pt_texture_render_lists.restype = i32
    L.pt_texture_render_lists.argtypes = [vp, sz, vp, sz, vp, sz, vp, i32, i32, i32, i32, vp,
                                          ctypes.POINTER(TextureStats)]
    L.pt_set_option.restype = i32
    L.pt_set_option.argtypes = [ctypes.c_char_p, i32]
    L.pt_get_option.restype = i32
    L.pt_get_option.argtypes = [ctypes.c_char_p, ctypes.POINTER(i32)]
    L.pt_kernel_launch_count.restype = ctypes.c_uint64
    _lib = L
    return L


_synth_lib = None


def synth_lib():
    """The synthetic-workload generators (bench / test scaffolding, libpt_synth_b200.so)."""
    global _synth_lib
    if _synth_lib is not None:
        return _synth_lib
    if not os.path.exists(SYNTH_LIB_PATH):
        raise ImportError(f"{SYNTH_LIB_PATH} is missing: run __graft_entry__.build()")
    L = ctypes.CDLL(SYNTH_LIB_PATH)
    vp, sz, i32, dbl = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double
    L.pt_synth_cloud_device.restype = i32
    L.pt_synth_cloud_device.argtypes = [vp, vp, sz, ctypes.POINTER(SynthParams), vp]
    L.pt_synth_samples_device.restype = i32
    L.pt_synth_samples_device.argtypes = [vp, sz, sz, dbl, dbl, dbl, dbl, i32, vp]
    L.pt_synth_pack_points_device.restype = i32
    L.pt_synth_pack_points_device.argtypes = [vp, vp, sz, vp, vp]
    L.pt_synth_pack_queries_device.restype = i32
    L.pt_synth_pack_queries_device.argtypes = [vp, sz, vp, vp]
    _synth_lib = L
    return L


def version():
    return lib().pt_version().decode()


def status_string(status):
    return lib().pt_status_string(int(status)).decode()


def device_count():
    return lib().pt_device_count()


def kernel_launch_count():
    return int(lib().pt_kernel_launch_count())


def debug_stats(reset=True):
    """Work counters of the query kernel (zeros unless the library was built with -DPT_STATS)."""
    buf = (ctypes.c_uint64 * 16)()
    _check(lib().pt_debug_stats(buf, 1 if reset else 0), "pt_debug_stats")
    names = ("expansions", "leaves", "pushes", "pops", "compactions", "heap_inserts", "parked",
             "warp_rounds", "overflowed", "samples", "warp_drain_iters", "warp_expand_iters",
             "grid_attempts", "grid_candidates", "grid_exact_selects", "grid_handed_over")
    return {n: int(buf[i]) for i, n in enumerate(names)}


def set_option(name, value):
    _check(lib().pt_set_option(name.encode(), int(value)), f"pt_set_option({name})")


def get_option(name):
    v = ctypes.c_int(0)
    _check(lib().pt_get_option(name.encode(), ctypes.byref(v)), f"pt_get_option({name})")
    return v.value


def _check(status, where):
    if status != PT_OK:
        raise PointsTransferError(status, where)


def make_points(xyz, normal=None, color=None, uv=None):
    """Pack arrays into the reference's 80-byte AoS ``Point`` records (src/Point.h)."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    p = np.zeros(xyz.shape[0], dtype=POINT_DTYPE)
    p["ver"] = xyz
    if normal is not None:
        p["normal"] = np.asarray(normal, dtype=np.float64).reshape(-1, 3)
    if color is not None:
        p["color"] = np.asarray(color, dtype=np.int32).reshape(-1, 3)
    if uv is not None:
        uv = np.asarray(uv, dtype=np.float64).reshape(-1, 2)
        p["U"], p["V"] = uv[:, 0], uv[:, 1]
    return p


def _as_points(a, what):
    a = np.ascontiguousarray(a)
    if a.dtype != POINT_DTYPE:
        raise TypeError(f"{what} must be an array of 80-byte Point records (POINT_DTYPE)")
    return a.reshape(-1)


def _radius(radius):
    return -1.0 if radius is None else float(radius)


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Distance:
    """The reference's metric (src/Distance.h).  Only the radius transform is needed on the
    host; the metric itself is evaluated on the device in the same operation order."""

    @staticmethod
    def transformed_distance(d):
        """``Distance::transformed_distance(double d)`` -- src/Distance.h:97."""
        return d * d

    @staticmethod
    def inverse_of_transformed_distance(d):
        """src/Distance.h:99."""
        return float(np.sqrt(d))


class Tree:
    """``Tree tree(points.begin(), points.end())`` -- src/pointsTransfer.cpp:259.

    ``points`` is an array of ``Point`` records on the host.  The spatial index is built
    on the GPU immediately (the reference builds lazily inside the first query)."""

    def __init__(self, points, device=-1, coord_mode=COORD_AUTO, ids=None):
        self._h = ctypes.c_void_p()
        pts = _as_points(points, "points")
        opts = BuildOpts()
        opts.device, opts.coord_mode = int(device), int(coord_mode)
        self._ids = None
        if ids is not None:
            self._ids = np.ascontiguousarray(ids, dtype=np.int32)
            if self._ids.shape[0] != pts.shape[0]:
                raise ValueError("ids must have one entry per point")
            opts.ids = self._ids.ctypes.data
        _check(lib().pt_index_build(_np_ptr(pts), pts.shape[0], ctypes.byref(opts),
                                    ctypes.byref(self._h)), "pt_index_build")

    @classmethod
    def _from_handle(cls, handle):
        t = cls.__new__(cls)
        t._h = handle
        t._ids = None
        return t

    @property
    def handle(self):
        return self._h

    def info(self):
        info = IndexInfo()
        _check(lib().pt_index_get_info(self._h, ctypes.byref(info)), "pt_index_get_info")
        return info

    def size(self):
        return int(self.info().n_points)

    def fallback_counts(self):
        """(samples re-run by the warp kernel, samples the grid kernel handed over) of the last
        query launch on this index."""
        out = (ctypes.c_uint32 * 2)()
        _check(lib().pt_index_fallback_counts(self._h, ctypes.byref(out)), "pt_index_fallback_counts")
        return int(out[0]), int(out[1])

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().pt_index_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- host-buffer calls (the reference-facing plugin surface) -----------------------
    def knn(self, queries, k, radius=None, want_d2=True, out_idx=None, out_d2=None):
        """Batched ``K_neighbor_search``: returns ``(idx[m,k] int32, d2[m,k] float64)``
        in ascending ``(d2, index)`` order, padded with ``-1`` / ``+inf``."""
        q = _as_points(queries, "queries")
        m = q.shape[0]
        idx = out_idx if out_idx is not None else np.empty((m, k), dtype=np.int32)
        d2 = out_d2 if out_d2 is not None else (np.empty((m, k), dtype=np.float64) if want_d2 else None)
        _check(lib().pt_knn(self._h, _np_ptr(q), m, int(k), _radius(radius), _np_ptr(idx),
                            _np_ptr(d2)), "pt_knn")
        return idx, d2

    def transfer(self, queries, k, radius=None, want_idx=True, want_d2=False, out=None):
        """k-NN + fused colour/normal blend for every sample.  Returns a dict with ``rgba``
        ``[m,4] uint8``, ``normal`` ``[m,3] float32`` and optionally ``idx`` / ``d2``."""
        q = _as_points(queries, "queries")
        m = q.shape[0]
        out = {} if out is None else out
        if "rgba" not in out:
            out["rgba"] = np.empty((m, 4), dtype=np.uint8)
        if "normal" not in out:
            out["normal"] = np.empty((m, 3), dtype=np.float32)
        if want_idx and "idx" not in out:
            out["idx"] = np.empty((m, k), dtype=np.int32)
        if want_d2 and "d2" not in out:
            out["d2"] = np.empty((m, k), dtype=np.float64)
        _check(lib().pt_transfer(self._h, _np_ptr(q), m, int(k), _radius(radius),
                                 _np_ptr(out.get("idx")), _np_ptr(out.get("d2")),
                                 _np_ptr(out["rgba"]), _np_ptr(out["normal"])), "pt_transfer")
        return out


    def texture(self, vertices, faces, k=20, resolution=8192, radius=None, pad=True):
        """The reference's output image (src/pointsTransfer.cpp:462-611): per-face transfer of the
        cloud's colours into the mesh's UV space.  vertices: Point records (position, U, V,
        colour), faces: int32 [F,3].  Returns (bgra uint8 [res,res,4] as cv::Mat CV_8UC4, stats)."""
        v = _as_points(vertices, "vertices")
        f = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        img = np.empty((resolution, resolution, 4), dtype=np.uint8)
        st = TextureStats()
        _check(lib().pt_texture_render(self._h, _np_ptr(v), v.shape[0], _np_ptr(f), f.shape[0], int(k),
                                       _radius(radius), int(resolution), 1 if pad else 0, _np_ptr(img),
                                       ctypes.byref(st)), "pt_texture_render")
        return img, {"triangles": int(st.triangles), "inside_points": int(st.inside_points),
                     "knn_ms": st.knn_ms, "draw_ms": st.draw_ms, "pad_ms": st.pad_ms}

    def transfer_slab(self, queries_ptr, queries_are_xyz, m, k, radius, boxes6, rank, halo,
                      idx_ptr, rgba_ptr, normal_ptr, d2_ptr=None):
        """pt_transfer_slab on raw host pointers (ints): one slab's pipelined host-buffer step
        with the ghost-zone check fused in.  boxes6: float64 numpy [R,6].  Returns True when the
        results are final (no sample may leave the ghost zone towards another slab)."""
        boxes6 = np.ascontiguousarray(boxes6, dtype=np.float64)
        need = ctypes.c_int32(0)
        _check(lib().pt_transfer_slab(self._h, queries_ptr, 1 if queries_are_xyz else 0, int(m),
                                      int(k), _radius(radius), _np_ptr(boxes6), boxes6.shape[0],
                                      int(rank), float(halo), idx_ptr, d2_ptr, rgba_ptr,
                                      normal_ptr, ctypes.byref(need)), "pt_transfer_slab")
        return need.value == 0


class ShardedTree:
    """The cloud sharded over several GPUs of one box by the C++ host (``pt_sharded_*``,
    csrc/pt_sharded.cu): same calls as ``Tree``, ids index the caller's ``points``."""

    def __init__(self, points, devices, halo=0.0, k_hint=20, coord_mode=COORD_AUTO):
        self._points = _as_points(points, "points")          # must outlive the handle
        self._h = ctypes.c_void_p()
        devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        o = ShardedOpts()
        o.n_devices, o.devices, o.halo, o.k_hint, o.coord_mode = len(devices), devs, float(halo), int(k_hint), int(coord_mode)
        _check(lib().pt_sharded_build(_np_ptr(self._points), self._points.shape[0], ctypes.byref(o),
                                      ctypes.byref(self._h)), "pt_sharded_build")

    def info(self):
        i = ShardedInfo()
        _check(lib().pt_sharded_get_info(self._h, ctypes.byref(i)), "pt_sharded_get_info")
        n = i.n_slabs
        return {"n_slabs": n, "rebuilds": i.rebuilds, "halo": i.halo, "slab_points": list(i.slab_points[:n]),
                "slab_ghosts": list(i.slab_ghosts[:n]), "slab_device": list(i.slab_device[:n])}

    def knn(self, queries, k, radius=None, want_d2=True):
        q = _as_points(queries, "queries")
        m = q.shape[0]
        idx = np.empty((m, k), dtype=np.int32)
        d2 = np.empty((m, k), dtype=np.float64) if want_d2 else None
        _check(lib().pt_sharded_knn(self._h, _np_ptr(q), m, int(k), _radius(radius), _np_ptr(idx), _np_ptr(d2)),
               "pt_sharded_knn")
        return idx, d2

    def transfer(self, queries, k, radius=None, want_idx=True, want_d2=False):
        q = _as_points(queries, "queries")
        m = q.shape[0]
        out = {"rgba": np.empty((m, 4), dtype=np.uint8), "normal": np.empty((m, 3), dtype=np.float32)}
        if want_idx:
            out["idx"] = np.empty((m, k), dtype=np.int32)
        if want_d2:
            out["d2"] = np.empty((m, k), dtype=np.float64)
        _check(lib().pt_sharded_transfer(self._h, _np_ptr(q), m, int(k), _radius(radius), _np_ptr(out.get("idx")),
                                         _np_ptr(out.get("d2")), _np_ptr(out["rgba"]), _np_ptr(out["normal"])),
               "pt_sharded_transfer")
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().pt_sharded_free(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def texture_from_lists(points, vertices, faces, idx, resolution=8192, pad=True, device=0):
    """``pt_texture_render_lists``: the texture stage for neighbour lists computed elsewhere."""
    p, v = _as_points(points, "points"), _as_points(vertices, "vertices")
    f = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    img = np.empty((resolution, resolution, 4), dtype=np.uint8)
    st = TextureStats()
    _check(lib().pt_texture_render_lists(_np_ptr(p), p.shape[0], _np_ptr(v), v.shape[0], _np_ptr(f), f.shape[0],
                                         _np_ptr(idx), idx.shape[1], int(resolution), 1 if pad else 0, int(device),
                                         _np_ptr(img), ctypes.byref(st)), "pt_texture_render_lists")
    return img, {"triangles": int(st.triangles), "inside_points": int(st.inside_points)}


class K_neighbor_search:
    """``K_neighbor_search search(tree, query, K)`` -- src/pointsTransfer.cpp:474.

    Iterating yields ``(point_index, squared_distance)`` in ascending order, like the
    reference's ``search.begin() .. search.end()`` over ``(Point, d2)`` pairs (:475-478);
    the caller indexes its own ``points`` array with ``point_index``."""

    def __init__(self, tree, query, k, radius=None):
        q = np.ascontiguousarray(query)
        if q.dtype != POINT_DTYPE:
            raise TypeError("query must be a Point record")
        idx, d2 = tree.knn(q.reshape(1), k, radius=radius)
        keep = idx[0] >= 0
        self._idx, self._d2 = idx[0][keep], d2[0][keep]

    def __iter__(self):
        return iter(zip(self._idx.tolist(), self._d2.tolist()))

    def __len__(self):
        return int(self._idx.shape[0])


# ---- device-buffer API (torch tensors as HBM handles; torch is plumbing only) ----------------

def _tptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


class DeviceTree(Tree):
    """Index built from coordinates already resident in HBM (bench `value`, slabs)."""

    def __init__(self, pos, attrs=None, ids=None):
        import torch
        assert pos.is_cuda and pos.is_contiguous() and pos.dim() == 2 and pos.shape[1] == 4
        assert pos.dtype in (torch.float32, torch.float64)
        n = pos.shape[0]
        if attrs is not None:
            assert attrs.is_cuda and attrs.is_contiguous() and attrs.dtype == torch.uint8
            assert attrs.numel() == 16 * n
        if ids is not None:
            assert ids.is_cuda and ids.is_contiguous() and ids.dtype == torch.int32
            assert ids.numel() == n
        self._h = ctypes.c_void_p()
        self._ids = None
        self.torch_device = pos.device
        torch.cuda.current_stream(pos.device).synchronize()
        _check(lib().pt_index_build_device(_tptr(pos), 1 if pos.dtype == torch.float64 else 0,
                                           _tptr(attrs), _tptr(ids), n, pos.device.index,
                                           ctypes.byref(self._h)), "pt_index_build_device")

    def query(self, queries, k, radius=None, radius2_per_query=None, idx=None, d2=None,
              rgba=None, normal=None, cand=None, stream=None):
        """Asynchronous on ``stream`` (default: torch's current stream).  ``queries`` is a
        float64 ``[m,3]`` CUDA tensor; outputs are pre-allocated CUDA tensors or None."""
        import torch
        assert queries.is_cuda and queries.dtype == torch.float64 and queries.is_contiguous()
        m = queries.shape[0]
        for t, dt, cnt in ((idx, torch.int32, m * k), (d2, torch.float64, m * k),
                           (rgba, torch.uint8, m * 4), (normal, torch.float32, m * 3),
                           (cand, torch.uint8, m * k * 32),
                           (radius2_per_query, torch.float64, m)):
            if t is not None:
                assert t.is_cuda and t.is_contiguous() and t.dtype == dt and t.numel() == cnt
        with torch.cuda.device(queries.device):
            _check(lib().pt_query_device(self._h, _tptr(queries), m, int(k), _radius(radius),
                                         _tptr(radius2_per_query), _tptr(idx), _tptr(d2),
                                         _tptr(rgba), _tptr(normal), _tptr(cand),
                                         _stream_ptr(stream)), "pt_query_device")


def merge_device(lists, n_lists, m, k, idx=None, d2=None, rgba=None, normal=None, cand=None,
                 stream=None):
    """K5: merge ``n_lists`` per-slab candidate lists (uint8 tensor of n_lists*m*k*32 bytes)."""
    import torch
    assert lists.is_cuda and lists.is_contiguous() and lists.dtype == torch.uint8
    assert lists.numel() == n_lists * m * k * 32
    with torch.cuda.device(lists.device):
        _check(lib().pt_merge_device(_tptr(lists), int(n_lists), int(m), int(k), _tptr(idx),
                                     _tptr(d2), _tptr(rgba), _tptr(normal), _tptr(cand),
                                     lists.device.index, _stream_ptr(stream)), "pt_merge_device")
