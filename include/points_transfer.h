/*
 * points_transfer.h -- C ABI of the B200-native pointsTransfer hot path.
 *
 * This is the drop-in boundary.  Each entry point replaces one call site of
 * the reference (horizon-research/3D-Reconstruction-From-Point-Cloud, paths
 * relative to its root):
 *
 *   pt_index_build   <- `Tree tree(points.begin(), points.end());`
 *                       src/pointsTransfer.cpp:259 (typedefs :37-40); the lazy
 *                       CGAL kd-tree build hidden in the first query as well.
 *   pt_knn           <- `K_neighbor_search search(tree, q, K);` + the result
 *                       iteration, src/pointsTransfer.cpp:474-478, metric
 *                       src/Distance.h:6-11, radius transform src/Distance.h:97.
 *   pt_transfer      <- the per-sample transfer loop src/pointsTransfer.cpp:
 *                       465-479 + the colour blend semantics of :95-103.
 *   pt_texture_render<- the face loop body after the searches and the texture
 *                       post-process, src/pointsTransfer.cpp:484-611 + draw_triangle :66-107.
 *   pt_index_free    <- `tree` going out of scope at src/pointsTransfer.cpp:626.
 *
 * Records are the reference's own 80-byte AoS `struct Point`
 * (src/Point.h:1-6): double ver[3] @0, double normal[3] @24, int color[3] @48,
 * double U @64, double V @72.  They are passed as `const void*` so a caller can
 * hand over `points.data()` of its `std::vector<Point>` unchanged.
 *
 * Contract: plain pointers and sizes only; int
 * status returned (0 = ok), nothing thrown across the ABI; blocking -- every
 * call is stream-synchronised on return; one handle <-> one caller thread at a
 * time (thread-compatible); results in query order; neighbours in ascending
 * (d2, index) order, ties broken by lowest point index; short lists padded with
 * index -1 / d2 +inf.  There is NO CPU fallback: without a CUDA device every
 * compute entry point returns PT_ERR_NO_DEVICE.
 */
#ifndef POINTS_TRANSFER_H
#define POINTS_TRANSFER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_OK                 0
#define PT_ERR_INVALID_ARG    1
#define PT_ERR_CUDA           2
#define PT_ERR_NO_DEVICE      3
#define PT_ERR_OUT_OF_MEMORY  4
#define PT_ERR_UNSUPPORTED    5  /* e.g. k > PT_MAX_K */
#define PT_ERR_NOT_REPRESENTABLE 6 /* coord_mode = PT_COORD_F32 but input is not fp32-exact */
#define PT_ERR_NON_FINITE     7  /* NaN/inf coordinate in the cloud */

#define PT_MAX_K 32
#define PT_POINT_STRIDE 80 /* sizeof(struct Point), src/Point.h */

/* Coordinate storage of the index.  The metric is always evaluated in fp64 in
 * the reference's operation order (src/Distance.h:6-11).  F32 storage (16-byte
 * float4 records) is lossless only when every cloud coordinate is exactly
 * representable in fp32; AUTO checks that on the device and falls back to the
 * 32-byte fp64 records otherwise, so neighbour indices stay bit-exact. */
#define PT_COORD_AUTO 0
#define PT_COORD_F32  1
#define PT_COORD_F64  2

typedef struct pt_index pt_index; /* opaque; owns device memory + one stream */

typedef struct pt_build_opts {
    int device;            /* CUDA ordinal; -1 = current device */
    int coord_mode;        /* PT_COORD_* */
    const int32_t *ids;    /* optional n global point ids, strictly increasing
                              (slab of a sharded cloud); NULL => 0..n-1 */
    int reserved[8];       /* must be zero */
} pt_build_opts;

typedef struct pt_index_info {
    uint64_t n_points;
    uint64_t n_leaves;       /* 32-point buckets of the curve-sorted cloud */
    int      n_levels;       /* box-pyramid levels */
    int      coord_mode;     /* PT_COORD_F32 or PT_COORD_F64 actually used */
    int      device;
    int      last_fallback_samples; /* samples the last query launch handed to the exact warp
                                       kernel (queue proof obligation not met); -1 unknown */
    double   bbox_lo[3], bbox_hi[3];
    uint64_t device_bytes;   /* resident after the build */
    float    build_ms;       /* device time of the last build (CUDA events) */
    float    last_query_ms;  /* device time of the last pt_knn/pt_transfer call: the chunked
                                H2D copy, kernels and D2H copy are pipelined over 3 streams */
    float    last_h2d_ms, last_d2h_ms; /* 0: overlapped, not separately measurable */
} pt_index_info;

/* Library / device probing (no compute). */
const char *pt_version(void);
const char *pt_status_string(int status);
int         pt_device_count(void); /* 0 when no CUDA device/driver */

/* Host-buffer API: the reference-facing plugin surface. ------------------- */

int pt_index_build(const void *points, size_t n, const pt_build_opts *opts,
                   pt_index **out);
int pt_index_free(pt_index *index);
int pt_index_get_info(const pt_index *index, pt_index_info *info);

/* Diagnostics of the last query launch on this index (synchronises the device):
 * out2[0] = samples re-run by the exact warp kernel (= last_fallback_samples),
 * out2[1] = samples the grid kernel handed to the box-pyramid kernels. */
int pt_index_fallback_counts(const pt_index *index, uint32_t out2[2]);

/* radius < 0 or +inf: unbounded k-NN (the reference's mode).  Otherwise only
 * points with d2 <= Distance::transformed_distance(radius) (src/Distance.h:97).
 * idx_out[m*k] (int32), d2_out[m*k] (double, may be NULL). */
int pt_knn(pt_index *index, const void *queries, size_t m, int k, double radius,
           int32_t *idx_out, double *d2_out);

/* k-NN + fused distance-weighted colour/normal blend (DESIGN.md "blend").
 * idx_out / d2_out may be NULL; rgba_out[m*4] (r,g,b,255), normal_out[m*3]. */
int pt_transfer(pt_index *index, const void *queries, size_t m, int k,
                double radius, int32_t *idx_out, double *d2_out,
                uint8_t *rgba_out, float *normal_out);

/* Everything the reference does after the neighbour search, for a whole mesh
 * (src/pointsTransfer.cpp:462-611): per face the union of the K nearest cloud points of its
 * corners, projection onto the face plane + in-triangle filter (:484-537), Delaunay
 * sub-triangulation with barycentric UVs (:539-581), rasterisation of every sub-triangle with
 * interpolated colours (draw_triangle, :66-107) and, with pad != 0, the 25x25 dilate + gutter
 * (:593-611).  vertices: n_vertices Point records (position, U, V, colour are read); faces:
 * 3 int32 vertex indices per face (n_faces < 2^24); bgra_out: resolution * resolution * 4 bytes
 * in HOST memory, row 0 first, byte order B G R A as cv::Mat CV_8UC4 -- what the reference hands
 * to cv::imwrite("texture.png").  Pixels nobody draws are 0 (the reference leaves them
 * uninitialised); the reference's out-of-bounds row / column writes are skipped. */
typedef struct pt_texture_stats {
    uint64_t triangles;      /* sub-triangles drawn */
    uint64_t inside_points;  /* neighbours kept by the in-triangle filter, all faces */
    float    knn_ms, draw_ms, pad_ms;   /* device time of the three stages */
} pt_texture_stats;
int pt_texture_render(pt_index *index, const void *vertices, size_t n_vertices,
                      const int32_t *faces, size_t n_faces, int k, double radius,
                      int resolution, int pad, uint8_t *bgra_out, pt_texture_stats *stats);

/* Device-buffer API: same operations with inputs/outputs already resident in
 * HBM on the index's device (bench `value`, multi-GPU slabs).  `stream` is a
 * cudaStream_t passed as void* (NULL = the CUDA default stream); these calls
 * are asynchronous with respect to the host unless stated. -------------- */

/* 16-byte point attribute record kept in original point order. */
typedef struct pt_attr {
    float   nx, ny, nz;
    uint8_t r, g, b, a;
} pt_attr;

/* 32-byte candidate record exchanged between slabs (NCCL payload). */
typedef struct pt_cand {
    double  d2;      /* +inf when empty */
    int32_t id;      /* global point id, -1 when empty */
    uint8_t r, g, b, a;
    float   nx, ny, nz;
    int32_t pad_;
} pt_cand;

/* pos: n records of 4 floats (x,y,z,unused) when coord_f64 == 0, or 4 doubles
 * (x,y,z,unused) when coord_f64 != 0.  attrs may be NULL (no blend possible).
 * ids as in pt_build_opts (device pointer).  Synchronises before returning. */
int pt_index_build_device(const void *pos, int coord_f64, const pt_attr *attrs,
                          const int32_t *ids, size_t n, int device,
                          pt_index **out);

/* queries_xyz: m*3 doubles.  Any output pointer may be NULL.  cand_out[m*k]
 * receives the per-slab candidate lists for the multi-GPU merge.
 * radius2_per_query (may be NULL): per-query squared search bound that
 * overrides `radius` (used for halo searches bounded by the owner's k-th d2;
 * a candidate must have d2 <= bound). */
int pt_query_device(pt_index *index, const double *queries_xyz, size_t m, int k,
                    double radius, const double *radius2_per_query,
                    int32_t *idx_out, double *d2_out, uint8_t *rgba_out,
                    float *normal_out, pt_cand *cand_out, void *stream);

/* K5: merge n_lists candidate lists per query (lists[l*m*k + q*k + j], each
 * ascending, padded) into the global top-k and blend.  Runs on `device`. */
int pt_merge_device(const pt_cand *lists, int n_lists, size_t m, int k,
                    int32_t *idx_out, double *d2_out, uint8_t *rgba_out,
                    float *normal_out, pt_cand *cand_out, int device, void *stream);

/* Halo exchange helpers for slab-sharded clouds.  Fixed-capacity routing,
 * so the exchange needs no host synchronisation: per peer a send block of (cap + 1) rows of 4
 * doubles -- row 0 = (count, overflow, 0, 0), rows 1.. = (x, y, z, squared bound) of the samples
 * whose k-th-neighbour ball reaches that peer's box -- and the sample index of every row.
 * boxes: n_ranks x 6 doubles (lo xyz, hi xyz).  counts: n_ranks words (reset by the call);
 * overflow_flag is OR-ed with 1 when a block did not fit (the caller then takes its exact
 * variable-size path). */
int pt_halo_route_device(const double *queries_xyz, const pt_cand *own_cand, size_t m, int k,
                         double radius, const double *boxes, int n_ranks, int self,
                         uint32_t cap, double *send, int32_t *sel, uint32_t *counts,
                         uint32_t *overflow_flag, void *stream);
/* received blocks (n_ranks x (cap+1) x 4) -> query rows + per-row squared bounds (-1 = unused) */
int pt_halo_prepare_device(const double *recv, int n_ranks, uint32_t cap, double *queries_out,
                           double *radius2_out, void *stream);
/* merge one peer's returned lists (cap x k records) into the owner's lists in place + re-blend */
int pt_halo_merge_device(pt_cand *own_cand, const pt_cand *back, const int32_t *sel,
                         const uint32_t *count, uint32_t cap, int k, int32_t *idx_out,
                         double *d2_out, uint8_t *rgba_out, float *normal_out, void *stream);

/* Sample routing of a slab-sharded cloud: samples arrive in arbitrary order; slab r owns the
 * samples with cuts[r] <= x < cuts[r + 1] (cuts: n_ranks + 1 doubles in DEVICE memory, cuts[0]
 * = -inf, cuts[n_ranks] = +inf).  Per destination a block of `cap` rows of 4 doubles (x, y, z,
 * +inf = per-query squared bound of pt_query_device); unused rows are NaN (answered with empty
 * lists), sel[r * cap + j] = index of the sample in that row or -1, counts[r] = rows used,
 * *overflow_flag = 1 when a block did not fit (nothing is dropped silently: enlarge cap). */
int pt_route_samples_device(const double *queries_xyz, size_t m, const double *cuts, int n_ranks,
                            uint32_t cap, double *send, int32_t *sel, uint32_t *counts,
                            uint32_t *overflow_flag, void *stream);
/* The way back: dst[sel[t]] = src[t] for every row t with sel[t] >= 0 (rows of row_bytes
 * bytes, a multiple of 4). */
int pt_scatter_rows_device(const void *src, const int32_t *sel, size_t rows, uint32_t row_bytes,
                           void *dst, void *stream);

/* Ghost-zone check for slab indexes that also hold the other slabs' points within `halo` of
 * this slab's box: ORs 1 into *flag if some sample's k-th-neighbour ball (d2[m*k], bounded by
 * radius) reaches another slab's box and may leave the ghost zone, i.e. the step needs the
 * exchange after all.  boxes as in pt_halo_route_device (the slabs' OWN boxes). */
int pt_ghost_check_device(const double *queries_xyz, const double *d2, size_t m, int k,
                          double radius, const double *boxes, int n_ranks, int self,
                          double halo, uint32_t *flag, void *stream);

/* Host-buffer step of ONE slab of a sharded cloud whose index also holds the other slabs' points
 * within `halo` of its box (ghost zone): pt_transfer on the samples this slab owns plus the
 * pt_ghost_check_device test, fused per pipeline chunk.  queries: m Point records (80 B), or
 * m x 3 doubles when queries_are_xyz != 0.  boxes: n_ranks x 6 doubles in HOST memory.
 * *needs_exchange = 1 when some sample's k-th-neighbour ball may leave the ghost zone towards
 * another slab: the results of such a call are not final and the caller runs the exchange
 * (DESIGN.md section 6).  Ids are the index's global ids. */
int pt_transfer_slab(pt_index *index, const void *queries, int queries_are_xyz, size_t m, int k,
                     double radius, const double *boxes, int n_ranks, int self, double halo,
                     int32_t *idx_out, double *d2_out, uint8_t *rgba_out, float *normal_out,
                     int *needs_exchange);

/* A cloud sharded over several GPUs of one box, host side in C++ (csrc/pt_sharded.cu): x-slabs
 * at point-count quantiles, one pt_index per slab holding the slab plus a ghost zone of width
 * `halo`, samples routed to the slab whose x-range holds them, one host thread per slab.  Same
 * contract as pt_knn / pt_transfer (blocking, query order, ascending (d2, index), ids are indices
 * into the caller's `points`); results are exact for any geometry: a call that finds a ghost
 * zone too narrow widens `halo`, rebuilds the slabs and answers again (counted in `rebuilds`).
 * `points` must stay valid until pt_sharded_free.  n < 2^31 points in total. */
typedef struct pt_sharded pt_sharded;
typedef struct pt_sharded_opts {
    int        n_devices;  /* number of slabs, 1..64 */
    const int *devices;    /* CUDA ordinal of every slab (may repeat); NULL => slab r on r mod device count */
    double     halo;       /* ghost-zone width; <= 0: 4x the expected k_hint-th neighbour distance */
    int        k_hint;     /* k the halo estimate is made for (default 20) */
    int        coord_mode; /* PT_COORD_* */
    int        reserved[8];
} pt_sharded_opts;
typedef struct pt_sharded_info {
    int      n_slabs, rebuilds;
    double   halo;
    uint64_t n_points;
    uint64_t slab_points[64], slab_ghosts[64];
    int      slab_device[64];
} pt_sharded_info;
int pt_sharded_build(const void *points, size_t n, const pt_sharded_opts *opts, pt_sharded **out);
int pt_sharded_free(pt_sharded *sharded);
int pt_sharded_get_info(const pt_sharded *sharded, pt_sharded_info *info);
int pt_sharded_knn(pt_sharded *sharded, const void *queries, size_t m, int k, double radius,
                   int32_t *idx_out, double *d2_out);
int pt_sharded_transfer(pt_sharded *sharded, const void *queries, size_t m, int k, double radius,
                        int32_t *idx_out, double *d2_out, uint8_t *rgba_out, float *normal_out);

/* pt_texture_render for neighbour lists that were computed elsewhere (e.g. pt_sharded_knn): the
 * cloud's positions and colours are uploaded to `device` in original order, idx[n_vertices * k]
 * (host) indexes `points`. */
int pt_texture_render_lists(const void *points, size_t n, const void *vertices, size_t n_vertices,
                            const int32_t *faces, size_t n_faces, const int32_t *idx, int k,
                            int resolution, int pad, int device, uint8_t *bgra_out,
                            pt_texture_stats *stats);

/* Tuning / introspection. */
/* Options: "knn_variant" (-1 auto [default]: grid kernel first, then scan / thread / warp for
 * the samples it hands over; 6 grid, 5 scan kernel, 2 thread kernel, 0 warp kernel),
 * "order" (0 Morton, 1 Hilbert [default], 2 Hilbert + kd refinement: no cell tables),
 * "grid" (1 build the uniform-grid cell tables [default]), "grid_tma" (1 stage candidate runs
 * with cp.async.bulk [default], 0 per-lane cp.async), "grid_pair" (1 [default]: two samples per
 * warp when k <= 16 and the first attempt's candidates are expected to fit a half-warp; 2: whenever
 * k <= 16; 0: never; the read-only "grid_pair_used" tells what the last launch did),
 * "grid_min_occ10" / "grid_admit100" / "grid_lookup_cost" (search-schedule tuning, see grid_plan:
 * finest table = finest level with >= 4.0 points per occupied cell; a block is tried first when
 * it reaches 1.00 expected k-th-neighbour distances; one bucket look-up costs 8 candidates), "sort_bits" (ordered key bits from the top, 8 per
 * radix pass; default 0 = auto: 40 up to 2^27 points, else 48),
 * "pool_keep_mb" (memory the library's private pool keeps mapped after a build or a free; default
 * -1 = up to a quarter of the device: a rebuild then reuses it -- 8.8 ms instead of 12-60 ms for
 * 50 M points -- and 0 hands everything back),
 * "sort" (1 hand-written radix sort [default], 0 cub), "host_chunks" (pipeline chunks of the
 * host-buffer API, default 8), "queue_cap" (tests: per-sample traversal queue entries, at most
 * the compiled 12), "verbose", "smem_pad" (diagnosis), "pool_guard" (debug: 256 guard bytes
 * around every device allocation of the library, compared when it is released; the read-only
 * "pool_guard_hits" counts the allocations found damaged). */
int pt_set_option(const char *name, int value);
int pt_get_option(const char *name, int *value);
/* Work counters of the query kernel since the last reset (16 words; all zero unless the library
 * was built with -DPT_STATS, a diagnosis build).  Synchronises the device. */
int pt_debug_stats(uint64_t *out16, int reset);
/* Number of kernels this library launched since process start (gpu_launches). */
uint64_t pt_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* POINTS_TRANSFER_H */
