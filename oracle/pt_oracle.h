/*
 * pt_oracle.h -- CPU oracle for the pointsTransfer detail-transfer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * product path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker (or as the timed CPU baseline), never as a fallback.
 *
 * PARITY OF THE SEARCH UNPINNED: the reference (horizon-research/3D-
 * Reconstruction-From-Point-Cloud) ships no tests, golden vectors or fixtures,
 * and its k-NN lives in CGAL Spatial_searching (find_package(CGAL), version
 * not pinned, src/CMakeLists.txt:12), which is absent from this image, so the
 * reference tool cannot be compiled or run here.  PINNED on the reference's
 * own headers (src/Point.h + src/Distance.h compiled by `make ref`,
 * ref_shim.cpp -> _ref/libref_metric.so; tests/test_oracle.py): the metric,
 * the box bound, new_distance, the radius transform and the record layout,
 * all bit for bit.  This file restates
 *   - the metric           src/Distance.h:6-11   (op order, fp64, no FMA)
 *   - the box lower bound  src/Distance.h:27-57
 *   - the incremental bound src/Distance.h:92-95
 *   - the radius transform src/Distance.h:97
 *   - the record layout    src/Point.h:1-6       (80-byte AoS Point)
 *   - the query call site  src/pointsTransfer.cpp:470-479
 * and CGAL's published algorithm (Kd_tree + Sliding_midpoint, bucket 10,
 * Orthogonal_k_neighbor_search with eps = 0), recalled from CGAL 4.14/5.x.
 * Exact k-NN with eps = 0 is implementation independent except at exact
 * distance ties, where this project imposes "lowest point index wins".
 * Independent cross-checks (scipy cKDTree, sklearn) live in tests/.
 */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Byte-for-byte mirror of the reference `struct Point` (src/Point.h:1-6):
 * double ver[3] @0, double normal[3] @24, int color[3] @48, (4 pad) @60,
 * double U @64, double V @72; sizeof == 80. */
typedef struct pto_point {
    double ver[3];
    double normal[3];
    int    color[3];
    int    pad_;
    double U;
    double V;
} pto_point;

/* src/Distance.h:6-11 -- squared Euclidean, (dx*dx + dy*dy) + dz*dz, fp64. */
double pto_transformed_distance(const pto_point *p1, const pto_point *p2);
/* src/Distance.h:27-57 -- point-to-box squared lower bound; writes per-axis
 * offsets into dists[] only when outside the slab (caller pre-zeroes). */
double pto_min_distance_to_rectangle(const pto_point *p, const double lo[3],
                                     const double hi[3], double dists[3]);
/* src/Distance.h:92-95 */
double pto_new_distance(double dist, double old_off, double new_off);
/* src/Distance.h:97 */
double pto_transformed_radius(double d);

/* Exact brute-force k-NN.  Ordering key (d2, index) lexicographic, lowest
 * index wins ties.  radius < 0 or +inf => unbounded; else only points with
 * d2 <= radius*radius.  Short lists padded with idx -1 / d2 +inf.
 * idx_out[m*k], d2_out[m*k] (d2_out may be NULL).  Returns 0. */
int pto_knn_bruteforce(const pto_point *pts, int64_t n, const pto_point *queries,
                       int64_t m, int k, double radius, int32_t *idx_out,
                       double *d2_out, int nthreads);

/* Frozen blend definition (DESIGN.md "blend"): inverse-squared-distance
 * weights, sequential fp64 summation in neighbour order, colour truncated
 * like src/pointsTransfer.cpp:100-102, alpha 255 like :103, normal
 * normalised and rounded to fp32.  rgba_out[m*4], normal_out[m*3]. */
int pto_blend(const pto_point *pts, int64_t n, int64_t m, int k,
              const int32_t *idx, const double *d2, uint8_t *rgba_out,
              float *normal_out);

/* CPU restatement of the reference's CPU path (CGAL Kd_tree, recalled):
 * sliding-midpoint splits, bucket_size points per leaf (CGAL default 10),
 * pointer-partitioned 80-byte records. */
typedef struct pto_kdtree pto_kdtree;
pto_kdtree *pto_kdtree_build(const pto_point *pts, int64_t n, int bucket_size);
void        pto_kdtree_free(pto_kdtree *t);
int64_t     pto_kdtree_node_count(const pto_kdtree *t);
/* exact_ties = 0: CGAL semantics (strict '<' against the current k-th, tie
 * order traversal dependent) -- the timing baseline.
 * exact_ties = 1: (d2, index) lexicographic, identical to the brute force. */
int pto_kdtree_knn(const pto_kdtree *t, const pto_point *queries, int64_t m,
                   int k, double radius, int exact_ties, int32_t *idx_out,
                   double *d2_out, int nthreads);

/* The reference's query set (src/pointsTransfer.cpp:465-479): for every face
 * j and corner i a K-NN at vertices[faces[j][i]].  Runs 3*F queries with the
 * CGAL-semantics search and returns the number of (point) results produced;
 * used only for the timed CPU baseline. */
int64_t pto_reference_face_loop(const pto_kdtree *t, const pto_point *vertices,
                                const int32_t *faces, int64_t face_count, int k,
                                int nthreads);

int pto_max_threads(void);

/* Per-face transfer + texture output (pt_texture_oracle.c): src/pointsTransfer.cpp:465-611 and
 * draw_triangle :66-107.  idx[n_vertices * k]: neighbour lists of the mesh vertices; bgra:
 * res * res * 4 bytes (B G R A), zeroed by the caller; stats[2] (may be NULL) += sub-triangles
 * drawn, inside points.  Sequential: later faces overwrite earlier ones, as in the reference. */
int pto_texture_faces(const pto_point *pts, const pto_point *vertices, const int32_t *idx, int k,
                      const int32_t *faces, int64_t n_faces, int res, uint8_t *bgra, int64_t *stats);
/* 25x25 dilate + gutter (src/pointsTransfer.cpp:593-611); in / out: res * res * 4 bytes */
int pto_texture_pad(const uint8_t *in, int res, uint8_t *out);

/* Host restatement of the synthetic workload generators (pt_synth_host.c): the clouds and mesh
 * samples of BASELINE.json's configs on the host cores, without any CUDA code (bench.py's
 * reference arm).  kind: 0 heightfield scan, 1 skewed clusters.  Returns 0. */
int pto_synth_cloud(pto_point *out, int64_t n, int kind, uint64_t seed, uint64_t first_index,
                    double u0, double u1, double v0, double v1, double sigma, int nthreads);
int pto_synth_samples(pto_point *out, int64_t gu, int64_t gv, double u0, double u1, double v0,
                      double v1, int center);

#ifdef __cplusplus
}
#endif
#endif
