// pt_grid.cu -- uniform-grid cell tables over the curve-sorted cloud (the "cell-start table" of
// SURVEY.md section 7 step 4; replaces the descent of the CGAL kd-tree behind
// `K_neighbor_search search(tree, q, K)`, /root/reference src/pointsTransfer.cpp:474, for samples
// whose neighbourhood has ordinary density -- the box pyramid stays the exact fallback).
//
// The sort key of a point interleaves the bits of its 21-bit lattice coordinates level by level
// (3 bits per level, Morton or Hilbert), so "the points of cell c at level l" is one contiguous
// run of the sorted cloud for every l.  For up to GRID_MAX_TABLES consecutive levels the build
// writes a hash table keyed by the PARENT cell (level l - 1) whose 32-byte bucket gives the run
// of each of the 8 children: start of the parent + cumulative child counts in curve order + the
// octant -> curve-rank permutation (Hilbert order rotates from cell to cell).  A 3x3x3 block of
// cells touches at most 8 buckets = 8 sectors.
//
//   grid_level_hist_kernel   cells per level: histogram of the coarsest level at which
//                            neighbouring sorted keys differ                       (1 read pass)
//   grid_insert_kernel       first point of every parent cell claims a bucket (CAS on the key)
//   grid_fill_kernel         first point of every child cell / last point of every parent cell
//                            writes the cumulative counts and the permutation     (no races:
//                            every 16-bit field has exactly one writer)
#include <cmath>
#include <cstring>

#include "pt_index.cuh"

namespace pt {

constexpr int GRID_BLOCK = 256;

// coarsest level (1..21) at which two keys fall into different cells; 22 when equal at `levels`
__device__ __forceinline__ int diff_level(unsigned long long a, unsigned long long b, int low_shift)
{
    const unsigned long long x = (a ^ b) >> low_shift;    // bits below the sorted prefix are noise
    if (x == 0) return 22;
    const int bit = 63 - __clzll((long long)x) + low_shift;    // 0..62
    return 21 - bit / 3;
}

__global__ void __launch_bounds__(GRID_BLOCK)
grid_level_hist_kernel(const unsigned long long *keys, uint32_t n, int low_shift,
                       unsigned long long *hist /* [24] */)
{
    __shared__ unsigned int h[24];
    if (threadIdx.x < 24) h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int ld = 23;
        if (i > 0) ld = diff_level(keys[i - 1], keys[i], low_shift);
        const unsigned peers = __match_any_sync(__activemask(), ld);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[ld], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 24 && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

struct GridBuildTables {
    GridBucket *buckets[GRID_MAX_TABLES];
    uint32_t    cap[GRID_MAX_TABLES];
    int         level[GRID_MAX_TABLES];
    int         n_tables;
    double      lo[3], inv_cell21;
    int         low_shift;
};

__device__ __forceinline__ uint32_t grid_hash(unsigned long long k)
{
    uint32_t h = (uint32_t)k * 0x9E3779B1u ^ (uint32_t)(k >> 32) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 12; h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}

template <typename PT>
__device__ __forceinline__ void lattice_of(const PT *pts, uint32_t i, const GridBuildTables &G,
                                           unsigned int c[3])
{
    const PT p = pts[i];
    const double v[3] = {(double)p.x, (double)p.y, (double)p.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {   // the same arithmetic as morton_kernel (pt_build.cu)
        double t = (v[a] - G.lo[a]) * G.inv_cell21;
        t = fmin(fmax(t, 0.0), 2097151.0);
        c[a] = (unsigned int)t;
    }
}

__device__ __forceinline__ unsigned long long parent_key(const unsigned int c[3], int child_level)
{
    const int sh = 22 - child_level;             // lattice -> parent level (child_level - 1)
    const unsigned long long px = sh >= 32 ? 0u : c[0] >> sh, py = sh >= 32 ? 0u : c[1] >> sh,
                             pz = sh >= 32 ? 0u : c[2] >> sh;
    return px | (py << 21) | (pz << 42);
}

template <typename PT>
__global__ void __launch_bounds__(GRID_BLOCK)
grid_insert_kernel(const unsigned long long *keys, const PT *pts, uint32_t n, GridBuildTables G)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ld = i ? diff_level(keys[i - 1], keys[i], G.low_shift) : 0;
    if (ld > G.level[0] - 1) return;             // not the first point of any parent cell
    unsigned int c[3];
    lattice_of(pts, i, G, c);
    for (int t = 0; t < G.n_tables; ++t) {
        if (ld > G.level[t] - 1) continue;
        const unsigned long long pk = parent_key(c, G.level[t]);
        // probe order: the two slots of the home group (one aligned 64-byte pair), then the
        // following groups -- the order grid_lookup (pt_knn_grid.cuh) reads them in
        uint32_t b = 2u * __umulhi(grid_hash(pk), G.cap[t] >> 1);
        for (;;) {
            GridBucket *B = G.buckets[t] + b;
            const unsigned long long old = atomicCAS(&B->key, ~0ull, pk);
            if (old == ~0ull) { B->start = i; B->perm = 0u; break; }
            if (++b == G.cap[t]) b = 0;
        }
    }
}

template <typename PT>
__global__ void __launch_bounds__(GRID_BLOCK)
grid_fill_kernel(const unsigned long long *keys, const PT *pts, uint32_t n, GridBuildTables G)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = keys[i];
    const unsigned long long kprev = i ? keys[i - 1] : 0ull;
    const int ld_prev = i ? diff_level(kprev, key, G.low_shift) : 0;
    const int ld_next = i + 1 < n ? diff_level(key, keys[i + 1], G.low_shift) : 0;
    if (ld_prev > G.level[0] && ld_next > G.level[0] - 1) return;
    unsigned int c[3];
    lattice_of(pts, i, G, c);
    for (int t = 0; t < G.n_tables; ++t) {
        const int L = G.level[t];
        const bool child_start = ld_prev <= L, parent_end = ld_next <= L - 1;
        if (!child_start && !parent_end) continue;
        const unsigned long long pk = parent_key(c, L);
        uint32_t b = 2u * __umulhi(grid_hash(pk), G.cap[t] >> 1);
        GridBucket *B;
        for (;;) {
            B = G.buckets[t] + b;
            if (B->key == pk) break;
            if (++b == G.cap[t]) b = 0;
        }
        const int ksh = 3 * (21 - L);
        const unsigned rank = (unsigned)(key >> ksh) & 7u;
        const uint32_t off = i - B->start;
        if (child_start) {
            const bool parent_start = ld_prev <= L - 1;
            const unsigned prev_rank = (unsigned)(kprev >> ksh) & 7u;
            const uint16_t o16 = (uint16_t)min(off, 0xffffu);
            for (unsigned r = parent_start ? 1u : prev_rank + 1u; r <= rank; ++r) B->cum[r] = o16;
            const int sh = 21 - L;
            const unsigned oct = ((c[0] >> sh) & 1u) | (((c[1] >> sh) & 1u) << 1) | (((c[2] >> sh) & 1u) << 2);
            atomicOr(&B->perm, (rank << (3 * oct)) | (1u << (24 + oct)));
        }
        if (parent_end) {
            const uint32_t total = off + 1;
            const uint16_t t16 = total > 0xfffeu ? (uint16_t)0xffffu : (uint16_t)total;
            B->cum[0] = t16;
            for (unsigned r = rank + 1; r < 8; ++r) B->cum[r] = t16;
        }
    }
}

static inline unsigned int cdiv_u(uint64_t a, uint64_t b) { return (unsigned int)((a + b - 1) / b); }

// Builds the cell tables of `ix` from the sorted keys (bits below `low_shift` unsorted) and the
// sorted point records.  Called by build_impl on ix->stream; synchronises once (level choice).
template <typename PT>
static int build_grid_impl(pt_index *ix, const unsigned long long *keys, int low_shift,
                           const double lo[3], double inv_cell21, double extent)
{
    cudaStream_t s = ix->stream;
    ix->grid = GridParams{};
    for (auto &c : ix->level_cells) c = 0;
    const uint32_t n = ix->n;
    if (n < 64 || !(inv_cell21 > 0.0) || opt_grid() == 0) return PT_OK;
    const int max_level = min(16, (63 - low_shift) / 3);     // levels whose cells are contiguous

    unsigned long long *d_hist = nullptr, h_hist[24];
    PT_TRY(pool_alloc((void **)&d_hist, sizeof h_hist, s));
    PT_CUDA(cudaMemsetAsync(d_hist, 0, sizeof h_hist, s));
    grid_level_hist_kernel<<<min(cdiv_u(n, GRID_BLOCK), 148u * 8u), GRID_BLOCK, 0, s>>>(keys, n, low_shift, d_hist);
    count_launch();
    PT_CUDA(cudaMemcpyAsync(h_hist, d_hist, sizeof h_hist, cudaMemcpyDeviceToHost, s));
    PT_CUDA(cudaStreamSynchronize(s));
    pool_free(d_hist, s);
    uint64_t cells = 1;
    ix->level_cells[0] = 1;
    for (int l = 1; l <= 21; ++l) {
        if (l <= max_level) cells += h_hist[l];
        ix->level_cells[l] = l <= max_level ? cells : 0;
    }
    // finest table: the finest level that still averages >= 4 points per occupied cell.  (A finer
    // one would only serve k < 8 a little better and costs most of the build: its bucket count
    // grows 4-8x per level.)
    int lf = 1;
    for (int l = 1; l <= max_level; ++l)
        if ((double)n / (double)ix->level_cells[l] >= 0.1 * (double)opt_grid_min_occ10()) lf = l;
    GridBuildTables G{};
    size_t total_buckets = 0;
    for (int t = 0; t < GRID_MAX_TABLES && lf - t >= 1; ++t) {
        const int L = lf - t;
        uint64_t cap = (3 * ix->level_cells[L - 1] + 8) & ~1ull;    // load <= 1/3, whole 2-slot groups
        if (cap > 0xfffffff0ull) break;
        G.level[t] = L;
        G.cap[t] = (uint32_t)cap;
        total_buckets += cap;
        G.n_tables = t + 1;
    }
    if (G.n_tables == 0) return PT_OK;
    GridBucket *mem = nullptr;
    PT_TRY(pool_alloc((void **)&mem, sizeof(GridBucket) * total_buckets, s));
    ix->grid_mem = mem;
    ix->grid_bytes = sizeof(GridBucket) * total_buckets;
    PT_CUDA(cudaMemsetAsync(mem, 0xff, ix->grid_bytes, s));
    for (int t = 0; t < G.n_tables; ++t) { G.buckets[t] = mem; mem += G.cap[t]; }
    for (int a = 0; a < 3; ++a) G.lo[a] = lo[a];
    G.inv_cell21 = inv_cell21;
    G.low_shift = low_shift;
    const PT *pts = reinterpret_cast<const PT *>(ix->pts);
    grid_insert_kernel<PT><<<cdiv_u(n, GRID_BLOCK), GRID_BLOCK, 0, s>>>(keys, pts, n, G);
    grid_fill_kernel<PT><<<cdiv_u(n, GRID_BLOCK), GRID_BLOCK, 0, s>>>(keys, pts, n, G);
    count_launch(2);
    PT_CUDA(cudaGetLastError());

    GridParams &gp = ix->grid;
    gp.n_tables = G.n_tables;
    for (int t = 0; t < G.n_tables; ++t) gp.tab[t] = GridTable{G.buckets[t], G.cap[t], G.level[t]};
    double amax = extent;
    for (int a = 0; a < 3; ++a) { gp.lo[a] = lo[a]; amax = fmax(amax, fabs(lo[a]) + extent); }
    gp.inv_cell21 = inv_cell21;
    gp.cell21 = extent / 2097152.0;
    // a point is assigned to a cell by fp64 arithmetic on its coordinates: distances to cell
    // boundaries are trusted only up to this margin (>> the rounding of (v - lo) * inv_cell21)
    gp.slack = 64.0 * 2.220446049250313e-16 * amax;
    return PT_OK;
}

int build_grid(pt_index *ix, const unsigned long long *keys, int low_shift, const double lo[3],
               double inv_cell21, double extent)
{
    return ix->coord_f64 ? build_grid_impl<PointD>(ix, keys, low_shift, lo, inv_cell21, extent)
                         : build_grid_impl<PointF>(ix, keys, low_shift, lo, inv_cell21, extent);
}

// Search schedule of one launch ("grid kernel").  occ(l) = mean points per
// occupied cell; the intrinsic dimension d of the cloud at that scale follows from how the cell
// count grows per level (a scanned surface: 2, a volume: 3); the expected k-th neighbour
// distance in cells is (k / (V_d occ))^(1/d).  An attempt (level, rc) is admissible when its
// guaranteed radius rc * cell covers that distance (or the radius bound); the cheapest
// admissible one -- fewest expected candidates -- goes first, then ever larger blocks.  (The
// block reaches rc + 0..1 cells beyond the sample, 1.125 on average for the nearest of six
// faces, so "rc >= expected distance" already succeeds for ~99 % of cfg4's samples; the 1.2x
// margin of the first version sent k = 32 on a 125 M-point slab to a level with 4x the
// candidates: 3.16 ms instead of 1.55 ms.  Options "grid_admit100", "grid_lookup_cost".)
int grid_plan(const pt_index *ix, int k, double r2, GridParams &gp)
{
    gp = ix->grid;
    gp.n_attempts = 0;
    if (gp.n_tables == 0) return 0;
    const double n = (double)ix->n;
    const double radius = r2 < INFINITY ? sqrt(r2) : INFINITY;
    int best_t = -1, best_rc = 0;
    double best_cost = INFINITY;
    for (int t = 0; t < gp.n_tables; ++t) {
        const int L = gp.tab[t].level;
        const double cells = (double)ix->level_cells[L], cells_up = (double)ix->level_cells[L - 1];
        const double occ = n / cells;
        double d = log2(fmax(cells / fmax(cells_up, 1.0), 1.0));
        d = fmin(fmax(d, 1.0), 3.0);
        const double vd = d <= 2.0 ? 2.0 + (d - 1.0) * (M_PI - 2.0) : M_PI + (d - 2.0) * (4.18879 - M_PI);
        double rk_cells = pow((double)k / (vd * occ), 1.0 / d);
        const double cell = gp.cell21 * (double)(1u << (21 - L));
        if (radius < INFINITY) rk_cells = fmin(rk_cells, radius / cell);
        for (int rc = 1; rc <= 2; ++rc) {
            if ((double)rc < 0.01 * (double)opt_grid_admit100() * rk_cells) continue;
            const double cost = pow(2.0 * rc + 1.0, d) * occ + (double)opt_grid_lookup_cost() * (rc == 1 ? 8 : 27);
            if (cost < best_cost) {
                best_cost = cost; best_t = t; best_rc = rc;
                gp.expect_cand = (float)(pow(2.0 * rc + 1.0, d) * occ);
            }
        }
    }
    if (best_t < 0) { best_t = gp.n_tables - 1; best_rc = 2; gp.expect_cand = 1e9f; }
    int t = best_t, rc = best_rc;
    while (gp.n_attempts < GRID_MAX_ATTEMPTS) {
        gp.att_tab[gp.n_attempts] = (unsigned char)t;
        gp.att_rc[gp.n_attempts] = (unsigned char)rc;
        ++gp.n_attempts;
        if (rc == 1) rc = 2;
        else if (t + 1 < gp.n_tables) ++t;
        else break;
    }
    return gp.n_attempts;
}

}  // namespace pt
