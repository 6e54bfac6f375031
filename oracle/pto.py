"""ctypes binding of the CPU oracle (oracle/libpt_oracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/pt_oracle.h.  Imported by tests/,
``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs; never by the product package.  The neighbour SEARCH is parity-unpinned (no reference
goldens, CGAL absent); the metric, the box bound and the record layout are pinned on the
reference's own headers through ``ref_metric()`` (oracle/ref_shim.cpp -> oracle/_ref/).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpt_oracle.so")

# Byte-for-byte mirror of the reference ``struct Point`` (src/Point.h:1-6).
POINT_DTYPE = np.dtype(
    [("ver", "<f8", 3), ("normal", "<f8", 3), ("color", "<i4", 3), ("pad_", "<i4"),
     ("U", "<f8"), ("V", "<f8")]
)
assert POINT_DTYPE.itemsize == 80


def build(force=False):
    """Compile the C restatement (gcc, oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f))
                                          for f in ("pt_oracle.c", "pt_synth_host.c", "pt_texture_oracle.c", "pt_oracle.h"))
    ):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_REF_PATH = os.path.join(_HERE, "_ref", "libref_metric.so")
_ref = None


def ref_metric():
    """The reference's OWN src/Point.h + src/Distance.h compiled by `make -C oracle ref`
    (oracle/ref_shim.cpp), or None where it was never built (/root/reference absent)."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_PATH):
            if os.path.exists("/root/reference/src/Distance.h"):
                subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True, stdout=subprocess.DEVNULL)
            else:
                return None
        R = ctypes.CDLL(_REF_PATH)
        vp, dbl = ctypes.c_void_p, ctypes.c_double
        R.ref_point_layout.argtypes = [vp]
        R.ref_transformed_distance.restype = dbl
        R.ref_transformed_distance.argtypes = [vp, vp]
        R.ref_min_distance_to_rectangle.restype = dbl
        R.ref_min_distance_to_rectangle.argtypes = [vp, vp, vp, vp]
        R.ref_new_distance.restype = dbl
        R.ref_new_distance.argtypes = [dbl, dbl, dbl]
        R.ref_transformed_radius.restype = dbl
        R.ref_transformed_radius.argtypes = [dbl]
        R.ref_inverse_of_transformed_distance.restype = dbl
        R.ref_inverse_of_transformed_distance.argtypes = [dbl]
        _ref = R
    return _ref


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i64, i32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
        L.pto_transformed_distance.restype = dbl
        L.pto_transformed_distance.argtypes = [vp, vp]
        L.pto_min_distance_to_rectangle.restype = dbl
        L.pto_min_distance_to_rectangle.argtypes = [vp, vp, vp, vp]
        L.pto_new_distance.restype = dbl
        L.pto_new_distance.argtypes = [dbl, dbl, dbl]
        L.pto_transformed_radius.restype = dbl
        L.pto_transformed_radius.argtypes = [dbl]
        L.pto_knn_bruteforce.restype = i32
        L.pto_knn_bruteforce.argtypes = [vp, i64, vp, i64, i32, dbl, vp, vp, i32]
        L.pto_blend.restype = i32
        L.pto_blend.argtypes = [vp, i64, i64, i32, vp, vp, vp, vp]
        L.pto_kdtree_build.restype = vp
        L.pto_kdtree_build.argtypes = [vp, i64, i32]
        L.pto_kdtree_free.restype = None
        L.pto_kdtree_free.argtypes = [vp]
        L.pto_kdtree_node_count.restype = i64
        L.pto_kdtree_node_count.argtypes = [vp]
        L.pto_kdtree_knn.restype = i32
        L.pto_kdtree_knn.argtypes = [vp, vp, i64, i32, dbl, i32, vp, vp, i32]
        L.pto_reference_face_loop.restype = i64
        L.pto_reference_face_loop.argtypes = [vp, vp, vp, i64, i32, i32]
        L.pto_max_threads.restype = i32
        u64 = ctypes.c_uint64
        L.pto_synth_cloud.restype = i32
        L.pto_synth_cloud.argtypes = [vp, i64, i32, u64, u64, dbl, dbl, dbl, dbl, dbl, i32]
        L.pto_synth_samples.restype = i32
        L.pto_synth_samples.argtypes = [vp, i64, i64, dbl, dbl, dbl, dbl, i32]
        L.pto_texture_faces.restype = i32
        L.pto_texture_faces.argtypes = [vp, vp, vp, i32, vp, i64, i32, vp, vp]
        L.pto_texture_pad.restype = i32
        L.pto_texture_pad.argtypes = [vp, i32, vp]
        _lib = L
    return _lib


def make_points(xyz, normal=None, color=None, uv=None):
    """Pack arrays into the 80-byte AoS ``Point`` records."""
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


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _chk_points(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == POINT_DTYPE, a.dtype
    return a


def transformed_distance(p1, p2):
    a, b = _chk_points(p1.reshape(1)), _chk_points(p2.reshape(1))
    return lib().pto_transformed_distance(_ptr(a), _ptr(b))


def min_distance_to_rectangle(p, lo, hi, dists=None):
    a = _chk_points(p.reshape(1))
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    d = np.zeros(3) if dists is None else dists
    return lib().pto_min_distance_to_rectangle(_ptr(a), _ptr(lo), _ptr(hi), _ptr(d)), d


def knn_bruteforce(points, queries, k, radius=-1.0, nthreads=0, want_d2=True):
    points, queries = _chk_points(points), _chk_points(queries)
    m = queries.shape[0]
    idx = np.empty((m, k), dtype=np.int32)
    d2 = np.empty((m, k), dtype=np.float64) if want_d2 else None
    rc = lib().pto_knn_bruteforce(_ptr(points), points.shape[0], _ptr(queries), m, k,
                                  float(radius), _ptr(idx), _ptr(d2) if want_d2 else None,
                                  nthreads)
    assert rc == 0, rc
    return idx, d2


def blend(points, idx, d2):
    points = _chk_points(points)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    d2 = np.ascontiguousarray(d2, dtype=np.float64)
    m, k = idx.shape
    rgba = np.empty((m, 4), dtype=np.uint8)
    nrm = np.empty((m, 3), dtype=np.float32)
    rc = lib().pto_blend(_ptr(points), points.shape[0], m, k, _ptr(idx), _ptr(d2),
                         _ptr(rgba), _ptr(nrm))
    assert rc == 0, rc
    return rgba, nrm


class KdTree:
    """The reference's CPU path restated: ``Tree tree(points.begin(), points.end())``
    (src/pointsTransfer.cpp:259) + ``K_neighbor_search`` (:474)."""

    def __init__(self, points, bucket_size=10):
        self._points = _chk_points(points)
        self._h = lib().pto_kdtree_build(_ptr(self._points), self._points.shape[0], bucket_size)
        assert self._h

    def close(self):
        if self._h:
            lib().pto_kdtree_free(self._h)
            self._h = None

    __del__ = close

    @property
    def node_count(self):
        return lib().pto_kdtree_node_count(self._h)

    def knn(self, queries, k, radius=-1.0, exact_ties=True, nthreads=0, want_d2=True):
        queries = _chk_points(queries)
        m = queries.shape[0]
        idx = np.empty((m, k), dtype=np.int32)
        d2 = np.empty((m, k), dtype=np.float64) if want_d2 else None
        rc = lib().pto_kdtree_knn(self._h, _ptr(queries), m, k, float(radius),
                                  1 if exact_ties else 0, _ptr(idx),
                                  _ptr(d2) if want_d2 else None, nthreads)
        assert rc == 0, rc
        return idx, d2

    def reference_face_loop(self, vertices, faces, k, nthreads=0):
        vertices = _chk_points(vertices)
        faces = np.ascontiguousarray(faces, dtype=np.int32)
        return lib().pto_reference_face_loop(self._h, _ptr(vertices), _ptr(faces),
                                             faces.shape[0], k, nthreads)


def max_threads():
    return lib().pto_max_threads()


def synth_cloud(n, seed, kind=0, u0=0.0, u1=1000.0, v0=0.0, v1=1000.0, sigma=0.01, first_index=0,
                nthreads=0):
    """Host restatement of the device cloud generator (pt_synth_host.c): ``n`` Point records."""
    out = np.empty(int(n), dtype=POINT_DTYPE)
    rc = lib().pto_synth_cloud(_ptr(out), int(n), int(kind), int(seed), int(first_index), u0, u1,
                               v0, v1, sigma, nthreads)
    assert rc == 0, rc
    return out


def synth_samples(gu, gv, u0=0.0, u1=1000.0, v0=0.0, v1=1000.0, center=False):
    """gu x gv mesh samples on the noise-free surface (row-major, v outer) as Point records."""
    out = np.empty(int(gu) * int(gv), dtype=POINT_DTYPE)
    rc = lib().pto_synth_samples(_ptr(out), int(gu), int(gv), u0, u1, v0, v1, 1 if center else 0)
    assert rc == 0, rc
    return out


def texture(points, vertices, idx, faces, resolution, pad=True):
    """The reference's face loop + texture post-process (pt_texture_oracle.c).  idx: int32
    [n_vertices, k] neighbour lists.  Returns (bgra uint8 [res,res,4], (triangles, inside))."""
    points, vertices = _chk_points(points), _chk_points(vertices)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
    img = np.zeros((resolution, resolution, 4), dtype=np.uint8)
    stats = np.zeros(2, dtype=np.int64)
    rc = lib().pto_texture_faces(_ptr(points), _ptr(vertices), _ptr(idx), idx.shape[1], _ptr(faces),
                                 faces.shape[0], resolution, _ptr(img), _ptr(stats))
    assert rc == 0, rc
    if pad:
        out = np.empty_like(img)
        assert lib().pto_texture_pad(_ptr(img), resolution, _ptr(out)) == 0
        img = out
    return img, (int(stats[0]), int(stats[1]))


def texture_pad(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.empty_like(img)
    assert lib().pto_texture_pad(_ptr(img), img.shape[0], _ptr(out)) == 0
    return out
