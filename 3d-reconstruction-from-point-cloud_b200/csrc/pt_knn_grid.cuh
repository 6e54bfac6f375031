// pt_knn_grid.cuh -- variant 6 ("grid"): one warp per sample over the uniform-grid cell tables
// (pt_grid.cu).  Replaces, for samples whose neighbourhood has ordinary density, the kd-tree
// descent of `K_neighbor_search search(tree, q, K)` (/root/reference src/pointsTransfer.cpp:474);
// the termination test is Distance::min_distance_to_rectangle (src/Distance.h:27-57) evaluated
// on the block of cells that was searched.
//
//   1. cell of the sample in O(1) from its lattice coordinates; lane c looks up cell c of the
//      (2 rc + 1)^3 block around it: one 32-byte bucket per parent cell (<= 8 sectors for 3^3)
//      gives (start, count) of the cell's run in the curve-sorted cloud;
//   2. a warp scan of the counts places the runs back to back in the warp's shared-memory
//      staging area and every lane that owns a run issues ONE bulk copy for it
//      (cp.async.bulk global -> shared, completion on an mbarrier; a run is contiguous and
//      16-byte aligned because records are float4 / 2 x double2);
//   3. lanes take the staged candidates round-robin: exact fp64 metric (src/Distance.h:6-11),
//      then a 32-bit selection key = fp32 bits of d2 rounded DOWN, low 8 bits replaced by the
//      candidate's slot.  Each lane sorts its <= 8 keys with a min/max network; the k + 1
//      smallest keys of the warp are extracted by k + 1 rounds of REDUX.MIN, the owner lane
//      popping its head.  Truncation is monotone, so distinct truncated keys order the exact
//      distances; if two neighbouring winners (or winner k and k + 1) share a truncated key the
//      sample is redone by an exact (d2, index) extraction;
//   4. final iff the k-th d2 (or the radius bound while the list is short) is <= the squared
//      distance from the sample to the boundary of the searched block (shrunk by a safety
//      margin); otherwise the next attempt of the launch's schedule searches a larger block,
//      and a sample that exhausts the schedule, meets a dense bucket or more than GRID_CAP
//      candidates is handed to the box-pyramid kernels (exact for any geometry);
//   5. winners' attributes gathered one per lane; the frozen blend's sequential fp64 sums run
//      one component per lane over a transposed shared-memory tile (bit-identical to BlendAcc).
//
// Two kernels share these steps: knn_grid_kernel (one sample per warp, any k <= 32, float4 or
// fp64 records, the whole attempt schedule) and knn_grid_pair_kernel (two samples per warp, 16
// lanes each, for k <= 16 on float4 records when the first attempt is a 3^3 block expected to
// stage fewer candidates than a half-warp holds; see the comment above it).  launch_grid picks.
#pragma once

namespace pt {

constexpr int GRID_CAP = 256;         // staged candidates per sample (8-bit slot numbers)
#ifndef PT_GRID_WARPS
#define PT_GRID_WARPS 4
#endif
constexpr int GRID_WARPS = PT_GRID_WARPS;
constexpr uint32_t GKEY_NONE = 0xffffffffu;

// ---- async-copy plumbing ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void ldgsts16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldgsts_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// ---- staged records ---------------------------------------------------------------------------
template <typename PT> struct GridRec;
template <> struct GridRec<PointF> {
    static __device__ __forceinline__ void load(const PointF *c, uint32_t j, double &x, double &y,
                                                double &z, int &idx)
    {
        const float4 v = *reinterpret_cast<const float4 *>(c + j);
        x = (double)v.x; y = (double)v.y; z = (double)v.z;
        idx = __float_as_int(v.w);
    }
};
template <> struct GridRec<PointD> {
    static __device__ __forceinline__ void load(const PointD *c, uint32_t j, double &x, double &y,
                                                double &z, int &idx)
    {
        const double2 *p = reinterpret_cast<const double2 *>(c + j);
        const double2 a = p[0], b = p[1];
        x = a.x; y = a.y; z = b.x;
        idx = __double2loint(b.y);
    }
};

__device__ __forceinline__ uint32_t grid_hash_q(unsigned long long k)   // = grid_hash of pt_grid.cu
{
    uint32_t h = (uint32_t)k * 0x9E3779B1u ^ (uint32_t)(k >> 32) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 12; h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}

// (start, count) of the run of cell (cx, cy, cz) at the table's level; count 0 when the cell is
// empty, 0xffffffff when its bucket is unusable (too many points for the 16-bit counts).
// Slots are probed two at a time (one aligned 64-byte group = one DRAM fetch): the two keys
// first, then the 24-byte payload of the slot that matched (an L1 hit).
__device__ __forceinline__ void grid_lookup(const GridBucket *buckets, uint32_t cap, uint32_t cx,
                                            uint32_t cy, uint32_t cz, uint32_t &start, uint32_t &cnt)
{
    const unsigned long long pkey = (unsigned long long)(cx >> 1) | ((unsigned long long)(cy >> 1) << 21) |
                                    ((unsigned long long)(cz >> 1) << 42);
    const unsigned oct = (cx & 1u) | ((cy & 1u) << 1) | ((cz & 1u) << 2);
    const uint32_t ngroups = cap >> 1;
    uint32_t g = __umulhi(grid_hash_q(pkey), ngroups);
    start = 0;
    cnt = 0;
    for (;;) {
        const unsigned long long *p = reinterpret_cast<const unsigned long long *>(buckets + 2 * (size_t)g);
        const unsigned long long k0 = __ldg(p), k1 = __ldg(p + 4);
        if (k0 == pkey || k1 == pkey) {
            const unsigned long long *b = k0 == pkey ? p : p + 4;
            const uint2 a = __ldg(reinterpret_cast<const uint2 *>(b + 1));     // start, perm
            if (((a.y >> (24 + oct)) & 1u) == 0u) return;
            const uint4 c = __ldg(reinterpret_cast<const uint4 *>(b + 2));     // cum[8]
            const unsigned rank = (a.y >> (3 * oct)) & 7u;
            const uint32_t total = c.x & 0xffffu;
            if (total == 0xffffu) { cnt = 0xffffffffu; return; }
            const unsigned r1 = rank + 1;
            // 16-bit field r of the 128-bit cum[]: pick the 64-bit half, then two bytes of it
            const uint32_t bx = rank < 4 ? c.x : c.z, by = rank < 4 ? c.y : c.w;
            const uint32_t ex = r1 < 4 ? c.x : c.z, ey = r1 < 4 ? c.y : c.w;
            const uint32_t beg = rank ? (__byte_perm(bx, by, 0x4410u + 0x22u * (rank & 3u)) & 0xffffu) : 0u;
            const uint32_t end = rank < 7 ? (__byte_perm(ex, ey, 0x4410u + 0x22u * (r1 & 3u)) & 0xffffu) : total;
            start = a.x + beg;
            cnt = end - beg;
            return;
        }
        if (k0 == ~0ull || k1 == ~0ull) return;
        if (++g == ngroups) g = 0;
    }
}

template <int C> __device__ __forceinline__ void gkey_sort(uint32_t (&h)[C]);
#define PT_GCAS(i, j) { const uint32_t lo_ = min(h[i], h[j]), hi_ = max(h[i], h[j]); h[i] = lo_; h[j] = hi_; }
template <> __device__ __forceinline__ void gkey_sort<1>(uint32_t (&)[1]) {}
template <> __device__ __forceinline__ void gkey_sort<2>(uint32_t (&h)[2]) { PT_GCAS(0, 1) }
template <> __device__ __forceinline__ void gkey_sort<4>(uint32_t (&h)[4])
{
    PT_GCAS(0, 1) PT_GCAS(2, 3) PT_GCAS(0, 2) PT_GCAS(1, 3) PT_GCAS(1, 2)
}
template <> __device__ __forceinline__ void gkey_sort<8>(uint32_t (&h)[8])   // Batcher, 19 exchanges
{
    PT_GCAS(0, 1) PT_GCAS(2, 3) PT_GCAS(4, 5) PT_GCAS(6, 7)
    PT_GCAS(0, 2) PT_GCAS(1, 3) PT_GCAS(4, 6) PT_GCAS(5, 7)
    PT_GCAS(1, 2) PT_GCAS(5, 6)
    PT_GCAS(0, 4) PT_GCAS(1, 5) PT_GCAS(2, 6) PT_GCAS(3, 7)
    PT_GCAS(2, 4) PT_GCAS(3, 5)
    PT_GCAS(1, 2) PT_GCAS(3, 4) PT_GCAS(5, 6)
}
template <> __device__ __forceinline__ void gkey_sort<12>(uint32_t (&h)[12])   // 39 exchanges, depth 9
{
    PT_GCAS(0, 8) PT_GCAS(1, 7) PT_GCAS(2, 6) PT_GCAS(3, 11) PT_GCAS(4, 10) PT_GCAS(5, 9)
    PT_GCAS(0, 1) PT_GCAS(2, 5) PT_GCAS(3, 4) PT_GCAS(6, 9) PT_GCAS(7, 8) PT_GCAS(10, 11)
    PT_GCAS(0, 2) PT_GCAS(1, 6) PT_GCAS(5, 10) PT_GCAS(9, 11)
    PT_GCAS(0, 3) PT_GCAS(1, 2) PT_GCAS(4, 6) PT_GCAS(5, 7) PT_GCAS(8, 11) PT_GCAS(9, 10)
    PT_GCAS(1, 4) PT_GCAS(3, 5) PT_GCAS(6, 8) PT_GCAS(7, 10)
    PT_GCAS(1, 3) PT_GCAS(2, 5) PT_GCAS(6, 9) PT_GCAS(8, 10)
    PT_GCAS(2, 3) PT_GCAS(4, 5) PT_GCAS(6, 7) PT_GCAS(8, 9)
    PT_GCAS(4, 6) PT_GCAS(5, 7)
    PT_GCAS(3, 4) PT_GCAS(5, 6) PT_GCAS(7, 8)
}
#undef PT_GCAS

// Fast selection on truncated keys.  On return lane r < k holds the key of the r-th smallest
// candidate (GKEY_NONE past the end) and `amb` tells whether some neighbouring pair of the first
// k + 1 keys shares its truncated distance (then the order / the cut is not proven).
// Slots [0, carry) and [32, total) of the staging area hold candidates (carry = 32 in the first
// round of a sample: everything below `total`; later rounds carry the previous winners in the
// first slots and stage the new window from slot 32 on).
template <typename PT, int C>
__device__ __forceinline__ void grid_select(const PT *cand, uint32_t carry, uint32_t total, double qx,
                                            double qy, double qz, double r2, int k, unsigned lane,
                                            uint32_t &mine, bool &amb)
{
    uint32_t h[C];
#pragma unroll
    for (int u = 0; u < C; ++u) {
        const uint32_t j = lane + 32u * u;
        uint32_t key = GKEY_NONE;
        if (j < total && (u > 0 || j < carry)) {
            double px, py, pz;
            int pidx;
            GridRec<PT>::load(cand, j, px, py, pz, pidx);
            const double d = dist2_exact(qx, qy, qz, px, py, pz);
            if (d <= r2) key = (__float_as_uint(__double2float_rd(d)) & ~0xffu) | j;
        }
        h[u] = key;
    }
    gkey_sort<C>(h);
    mine = GKEY_NONE;
    uint32_t m = GKEY_NONE;
    // k + 1 rounds, no early exit: once the candidates run out every round yields GKEY_NONE
#pragma unroll 2
    for (int r = 0; r <= k; ++r) {
        m = __reduce_min_sync(0xffffffffu, h[0]);
        if ((unsigned)r == lane) mine = m;
        if ((m & 31u) == lane) {           // the owner pops its head (a GKEY_NONE round pops a GKEY_NONE)
#pragma unroll
            for (int u = 0; u + 1 < C; ++u) h[u] = h[u + 1];
            h[C - 1] = GKEY_NONE;
        }
    }
    uint32_t nxt = __shfl_down_sync(0xffffffffu, mine, 1);
    if (lane == 31) nxt = m;                    // k == 32: the 33rd key is only in m
    amb = __any_sync(0xffffffffu, lane < (unsigned)k && mine != GKEY_NONE && (mine >> 8) == (nxt >> 8));
    if (lane >= (unsigned)k) mine = GKEY_NONE;
}

// Exact selection on (d2, index): k rounds of a three-stage warp argmin.  Slow, rare (equal
// truncated keys: ~0.5 % of the samples of a scanned surface, every sample of a lattice).
template <typename PT>
__device__ __forceinline__ void grid_select_exact(const PT *cand, uint32_t carry, uint32_t total,
                                                  double qx, double qy, double qz, double r2, int k,
                                                  unsigned lane, uint32_t &mine)
{
    uint32_t taken = 0;
    mine = GKEY_NONE;
#pragma unroll 1
    for (int r = 0; r < k; ++r) {
        double bd = INFINITY;
        int bi = IDX_NONE, bu = -1;
#pragma unroll 1
        for (int u = 0; lane + 32u * u < total; ++u) {
            if (((taken >> u) & 1u) || (u == 0 && lane >= carry)) continue;
            double px, py, pz;
            int pidx;
            GridRec<PT>::load(cand, lane + 32u * u, px, py, pz, pidx);
            const double d = dist2_exact(qx, qy, qz, px, py, pz);
            if (d <= r2 && (bu < 0 || key_less(d, pidx, bd, bi))) { bd = d; bi = pidx; bu = u; }
        }
        bool in = bu >= 0;
        if (!__any_sync(0xffffffffu, in)) break;
        // every lane takes part in every reduction (no short-circuit around a *_sync call)
        const uint32_t hi = in ? (uint32_t)__double2hiint(bd) : 0xffffffffu;
        const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
        in = in && hi == mh;
        const uint32_t lo = in ? (uint32_t)__double2loint(bd) : 0xffffffffu;
        const uint32_t ml = __reduce_min_sync(0xffffffffu, lo);
        in = in && lo == ml;
        const uint32_t ii = in ? (uint32_t)bi : 0xffffffffu;
        const uint32_t mi = __reduce_min_sync(0xffffffffu, ii);
        in = in && ii == mi;
        const int w = __ffs(__ballot_sync(0xffffffffu, in)) - 1;
        const uint32_t jw = __shfl_sync(0xffffffffu, (uint32_t)(lane + 32u * (bu < 0 ? 0 : bu)), w);
        if (lane == (unsigned)w) taken |= 1u << bu;
        if ((unsigned)r == lane) mine = jw;
    }
}

#ifdef PT_STATS
#define PT_GSTAT(slot, v) do { if (lane == 0) atomicAdd(&g_stats[slot], (unsigned long long)(v)); } while (0)
#else
#define PT_GSTAT(slot, v) ((void)0)
#endif
// stats slots of the grid kernel: 12 attempts, 13 candidates staged, 14 exact selections,
// 15 samples handed to the box-pyramid kernels

constexpr int GRID_MAX_ROUNDS = 8;    // staging rounds per attempt: 256 + 7 * 224 candidates

// One attempt: look up the (2 RC + 1)^3 block of cells around cell c, stage its points -- in
// several rounds if they exceed the staging area, the winners so far carried in slots [0, k) --
// and select the k best.  Returns 0 (lane r's `mine` = slot key of the r-th best, GKEY_NONE past
// the end), 1 (the block cannot fill the list: try the next attempt) or 2 (hand the sample over).
template <typename PT, int RC, bool TMA>
__device__ __forceinline__ int grid_attempt(const QueryParams &P, const GridTable &T, const uint32_t c[3],
                                            double qx, double qy, double qz, double r2, bool short_ok,
                                            unsigned lane, PT *cand, uint32_t bar, uint32_t &phase,
                                            uint32_t &mine)
{
    constexpr int SIDE = 2 * RC + 1, CELLS = SIDE * SIDE * SIDE, NP = (CELLS + 31) / 32;
    const int k = P.k;
    const uint32_t ncell = 1u << T.level;
    uint32_t st[NP], ct[NP], mysum = 0;
    bool dense = false;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int nidx = 32 * p + (int)lane;
        st[p] = 0; ct[p] = 0;
        if (nidx < CELLS) {
            const uint32_t x = c[0] + (uint32_t)(nidx % SIDE) - RC;
            const uint32_t y = c[1] + (uint32_t)((nidx / SIDE) % SIDE) - RC;
            const uint32_t z = c[2] + (uint32_t)(nidx / (SIDE * SIDE)) - RC;
            if (x < ncell && y < ncell && z < ncell)       // unsigned: also rejects "negative" cells
                grid_lookup(T.buckets, T.cap, x, y, z, st[p], ct[p]);
            if (ct[p] == 0xffffffffu) { dense = true; ct[p] = 0; }
            mysum += ct[p];
        }
    }
    uint32_t incl = mysum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += y;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (__any_sync(0xffffffffu, dense)) return 2;
    PT_GSTAT(13, total);
    if (total < (uint32_t)k && !short_ok) return 1;
    if (total > (uint32_t)(GRID_CAP + (GRID_MAX_ROUNDS - 1) * (GRID_CAP - 32))) return 2;
    const uint32_t excl = incl - mysum;       // this lane's runs cover candidates [excl, excl + mysum)
    const PT *pts = reinterpret_cast<const PT *>(P.pts);
    mine = GKEY_NONE;
    uint32_t w0 = 0, carry = 32;
    while (w0 < total) {
        const uint32_t base = w0 == 0 ? 0u : 32u;
        const uint32_t w1 = min(total, w0 + (uint32_t)GRID_CAP - base);
        // ---- stage candidates [w0, w1) into slots [base, base + w1 - w0) --------------------------
        if (TMA) {
            // the staging area was read (and the blend tile written) through the generic proxy
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_expect(bar, (w1 - w0) * (uint32_t)sizeof(PT));
            __syncwarp();
            uint32_t o = excl;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t lo = max(o, w0), hi = min(o + ct[p], w1);
                if (lo < hi)
                    bulk_g2s(smem_u32(cand + base + (lo - w0)), pts + st[p] + (lo - o),
                             (hi - lo) * (uint32_t)sizeof(PT), bar);
                o += ct[p];
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            __syncwarp();
            uint32_t o = excl;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t lo = max(o, w0), hi = min(o + ct[p], w1);
                unsigned msk = __ballot_sync(0xffffffffu, lo < hi);
                while (msk) {
                    const int e = __ffs(msk) - 1;
                    msk &= msk - 1;
                    const uint32_t s0 = __shfl_sync(0xffffffffu, st[p] + (lo - o), e);
                    const uint32_t c0 = __shfl_sync(0xffffffffu, hi - lo, e);
                    const uint32_t d0 = __shfl_sync(0xffffffffu, base + (lo - w0), e);
                    for (uint32_t i = lane; i < c0 * (uint32_t)(sizeof(PT) / 16); i += 32)
                        ldgsts16(smem_u32(cand + d0) + 16u * i, reinterpret_cast<const char *>(pts + s0) + 16u * i);
                }
                o += ct[p];
            }
            ldgsts_wait_all();
            __syncwarp();
        }
        // ---- select -----------------------------------------------------------------------------
        const uint32_t slots = base + (w1 - w0);
        bool amb;
        if (slots <= 64u) grid_select<PT, 2>(cand, carry, slots, qx, qy, qz, r2, k, lane, mine, amb);
        else if (slots <= 128u) grid_select<PT, 4>(cand, carry, slots, qx, qy, qz, r2, k, lane, mine, amb);
        else grid_select<PT, 8>(cand, carry, slots, qx, qy, qz, r2, k, lane, mine, amb);
        if (amb) {
            PT_GSTAT(14, 1);
            grid_select_exact<PT>(cand, carry, slots, qx, qy, qz, r2, k, lane, mine);
        }
        w0 = w1;
        if (w0 < total) {       // more to come: the winners so far move to slots [0, count)
            PT rec{};
            if (mine != GKEY_NONE) rec = cand[mine & 0xffu];
            __syncwarp();
            if (mine != GKEY_NONE) cand[lane] = rec;
            carry = (uint32_t)__popc(__ballot_sync(0xffffffffu, mine != GKEY_NONE));
            __syncwarp();
        }
    }
    return 0;
}

template <typename PT, bool TMA>
__device__ __forceinline__ void grid_sample(const QueryParams &P, uint32_t s, double qx, double qy,
                                            double qz, unsigned lane, PT *cand, uint32_t bar,
                                            uint32_t &phase, uint32_t *ovf_count, uint32_t *ovf_list)
{
    const int k = P.k;
    const double r2 = P.r2_per_query ? __ldg(P.r2_per_query + s) : P.r2;
    const GridParams &G = P.grid;
    // lattice coordinates: the arithmetic of morton_kernel (pt_build.cu); the conversion
    // saturates, so negative / NaN -> 0 exactly like its fmax(t, 0)
    const double t[3] = {(qx - G.lo[0]) * G.inv_cell21, (qy - G.lo[1]) * G.inv_cell21,
                         (qz - G.lo[2]) * G.inv_cell21};
    uint32_t c21[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) c21[a] = min(__double2uint_rz(t[a]), 2097151u);
    const double slack = G.slack + 1.8e-15 * fmax(fmax(fabs(qx), fabs(qy)), fabs(qz));

    uint32_t mine = GKEY_NONE;
    double d = INFINITY;                // exact d2 of this lane's neighbour
    int li = IDX_NONE;
    bool done = !(r2 >= 0.0);           // a negative bound admits nothing: the empty answer is final
    for (int a = 0; a < G.n_attempts && !done; ++a) {
        const GridTable T = G.tab[G.att_tab[a]];
        const int rc = G.att_rc[a];
        const int sh = 21 - T.level;
        const uint32_t c[3] = {c21[0] >> sh, c21[1] >> sh, c21[2] >> sh};
        // Distance from the sample to the nearest face of the searched block that has cells
        // beyond it (src/Distance.h:27-57 on the block, seen from inside), in lattice units:
        // position inside the own cell, bracketed in fp32 and combined with directed rounding so
        // that the result never exceeds the true distance.
        const uint32_t ncell = 1u << T.level;
        const float S = (float)(1u << sh), below = (float)rc * S, above = (float)(rc + 1) * S;
        float gl = INFINITY;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            const double f = t[ax] - (double)(c[ax] << sh);
            const float dl = __fadd_rd(__double2float_rd(f), below);
            const float du = __fsub_rd(above, __double2float_ru(f));
            gl = fminf(gl, c[ax] > (uint32_t)rc ? dl : INFINITY);
            gl = fminf(gl, c[ax] + rc + 1 < ncell ? du : INFINITY);
        }
        const double g = (double)gl * G.cell21 - slack;
        const double g2 = g > 0.0 ? __dmul_rd(g, g) : 0.0;
        PT_GSTAT(12, 1);
        int rcode;
        if (rc == 1) rcode = grid_attempt<PT, 1, TMA>(P, T, c, qx, qy, qz, r2, r2 <= g2, lane, cand, bar, phase, mine);
        else rcode = grid_attempt<PT, 2, TMA>(P, T, c, qx, qy, qz, r2, r2 <= g2, lane, cand, bar, phase, mine);
        if (rcode == 2) break;                     // dense bucket / too many candidates
        if (rcode == 1) continue;                  // cannot fill the list and cannot prove a short one
        d = INFINITY;
        li = IDX_NONE;
        if (mine != GKEY_NONE) {
            double px, py, pz;
            GridRec<PT>::load(cand, mine & 0xffu, px, py, pz, li);
            d = dist2_exact(qx, qy, qz, px, py, pz);
        }
        const double kth = __shfl_sync(0xffffffffu, d, k - 1);     // +inf while the list is short
        done = fmin(kth, r2) <= g2;
    }
    if (!done) {
        PT_GSTAT(15, 1);
        if (lane == 0) ovf_list[atomicAdd(ovf_count, 1u)] = s;
        return;
    }

    // ---- outputs: lane r holds the r-th neighbour ------------------------------------------------
    const bool has = mine != GKEY_NONE;
    const int gid = has ? (P.ids ? __ldg(P.ids + li) : li) : -1;
    const bool want_blend = P.rgba_out || P.normal_out;
    AttrRaw at{0.f, 0.f, 0.f, 0u};
    if (has && (want_blend || P.cand_out) && P.attrs) at = load_attr(P.attrs + li);
    if (lane < (unsigned)k) {
        const size_t o = (size_t)s * k + lane;
        if (P.idx_out) P.idx_out[o] = gid;
        if (P.d2_out) P.d2_out[o] = d;
        if (P.cand_out) store_cand(P.cand_out + o, d, gid, at);
    }
    if (!want_blend) return;
    uint8_t *ro = P.rgba_out ? P.rgba_out + 4 * (size_t)s : nullptr;
    float *no = P.normal_out ? P.normal_out + 3 * (size_t)s : nullptr;
    const int cnt = __popc(__ballot_sync(0xffffffffu, has));
    if (cnt == 0) {
        if (lane == 0) store_empty_blend(ro, no);
        return;
    }
    // frozen blend: per-neighbour terms in parallel, the sequential sums
    // one component per lane over a [7][32] tile that reuses the staging area
    const int mode = __shfl_sync(0xffffffffu, d, 0) == 0.0 ? 1 : 0;
    const double w = has ? (mode ? (d == 0.0 ? 1.0 : 0.0) : __ddiv_rn(1.0, d)) : 0.0;
    double *tile = reinterpret_cast<double *>(cand);
    __syncwarp();                               // every lane has read its winner's record
    tile[0 * 32 + lane] = w;                    // lanes past the list write zeros that are never read
    tile[1 * 32 + lane] = __dmul_rn(w, (double)(at.rgba & 0xffu));
    tile[2 * 32 + lane] = __dmul_rn(w, (double)((at.rgba >> 8) & 0xffu));
    tile[3 * 32 + lane] = __dmul_rn(w, (double)((at.rgba >> 16) & 0xffu));
    tile[4 * 32 + lane] = __dmul_rn(w, (double)at.nx);
    tile[5 * 32 + lane] = __dmul_rn(w, (double)at.ny);
    tile[6 * 32 + lane] = __dmul_rn(w, (double)at.nz);
    __syncwarp();
    double acc = 0.0;
    if (lane < 7) {
        const double *row = tile + lane * 32;
        int j = 0;
#pragma unroll 1
        for (; j + 4 <= cnt; j += 4) {
            acc = __dadd_rn(acc, row[j]);
            acc = __dadd_rn(acc, row[j + 1]);
            acc = __dadd_rn(acc, row[j + 2]);
            acc = __dadd_rn(acc, row[j + 3]);
        }
#pragma unroll 1
        for (; j < cnt; ++j) acc = __dadd_rn(acc, row[j]);
    }
    double s0 = __shfl_sync(0xffffffffu, acc, 0);
    if (!(s0 > 0.0 && s0 < INFINITY)) {         // overflowed weights: nearest neighbour only
        const uint32_t c0 = __shfl_sync(0xffffffffu, at.rgba, 0);
        const float n0x = __shfl_sync(0xffffffffu, at.nx, 0), n0y = __shfl_sync(0xffffffffu, at.ny, 0),
                    n0z = __shfl_sync(0xffffffffu, at.nz, 0);
        const double v[7] = {1.0, (double)(c0 & 0xffu), (double)((c0 >> 8) & 0xffu), (double)((c0 >> 16) & 0xffu),
                             (double)n0x, (double)n0y, (double)n0z};
        acc = 0.0;
#pragma unroll
        for (int a = 0; a < 7; ++a) if (lane == (unsigned)a) acc = v[a];
        s0 = 1.0;
    }
    const double s4 = __shfl_sync(0xffffffffu, acc, 4), s5 = __shfl_sync(0xffffffffu, acc, 5),
                 s6 = __shfl_sync(0xffffffffu, acc, 6);
    const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(s4, s4), __dmul_rn(s5, s5)), __dmul_rn(s6, s6)));
    const double den = lane < 4 ? s0 : len;
    const double quo = __ddiv_rn(acc, den);
    const int ci = min(max(__double2int_rz(quo), 0), 255);
    const int cr = __shfl_sync(0xffffffffu, ci, 1), cg = __shfl_sync(0xffffffffu, ci, 2),
              cb = __shfl_sync(0xffffffffu, ci, 3);
    if (lane == 0 && ro) *reinterpret_cast<uchar4 *>(ro) = make_uchar4((unsigned char)cr, (unsigned char)cg, (unsigned char)cb, 255);
    if (lane >= 4 && lane < 7 && no)
        no[lane - 4] = (len > 0.0 && len < INFINITY) ? __double2float_rn(quo) : 0.0f;
}

#ifndef PT_GRID_MIN_BLOCKS
#define PT_GRID_MIN_BLOCKS 8
#endif
template <typename PT, bool TMA>
__global__ void __launch_bounds__(GRID_WARPS * 32, PT_GRID_MIN_BLOCKS)
knn_grid_kernel(const QueryParams P, uint32_t *ovf_count, uint32_t *ovf_list)
{
    extern __shared__ __align__(128) unsigned char g_smem[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    PT *cand = reinterpret_cast<PT *>(g_smem) + (size_t)wib * GRID_CAP;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(g_smem + sizeof(PT) * GRID_CAP * GRID_WARPS);
    const uint32_t bar = smem_u32(bars + wib);
    if (TMA) {
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
    }
    uint32_t phase = 0;
    const uint32_t n_warps = gridDim.x * GRID_WARPS;
    uint32_t s = blockIdx.x * GRID_WARPS + wib;
    if (s >= P.m) return;
    double qx = __ldg(P.queries + 3 * (size_t)s), qy = __ldg(P.queries + 3 * (size_t)s + 1),
           qz = __ldg(P.queries + 3 * (size_t)s + 2);
    while (s < P.m) {
        // the next sample's coordinates are fetched while this one is answered
        const uint32_t sn = s + n_warps;
        double nx = 0.0, ny = 0.0, nz = 0.0;
        if (sn < P.m) {
            nx = __ldg(P.queries + 3 * (size_t)sn);
            ny = __ldg(P.queries + 3 * (size_t)sn + 1);
            nz = __ldg(P.queries + 3 * (size_t)sn + 2);
        }
        grid_sample<PT, TMA>(P, s, qx, qy, qz, lane, cand, bar, phase, ovf_count, ovf_list);
        s = sn; qx = nx; qy = ny; qz = nz;
    }
}


// ---- two samples per warp (k <= 16, float4 records, first attempt with rc = 1) ------------------
// Most of a sample's instructions are warp-uniform bookkeeping that 32 lanes execute for one
// sample (set-up, scan, selection rounds, blend epilogue: profiles/README.md).  Here a warp
// answers samples 2p and 2p + 1 at once, 16 lanes each: sub-lane c looks up cells c and c + 16 of
// the 3^3 block, every half stages its runs into its own PAIR_CAP slots (one mbarrier for
// both), lanes take the staged candidates 16 apart, the selection rounds reduce over 16-lane
// segments, and sub-lane r ends up with the r-th neighbour.  A half that cannot be proven final
// by this first attempt (list not full, more than PAIR_CAP candidates, dense bucket, k-th
// distance beyond the block) is redone by the whole warp through grid_sample -- the complete
// attempt schedule, staging rounds and the hand-over to the box-pyramid kernels.
// (Two neighbours per sub-lane for 16 < k <= 32 was built and measured: exact, but 1.74 ms against
// the one-sample kernel's 1.55 ms at k = 32 on a 125 M-point slab of cfg4 and 1.19 against 1.21 ms
// at k = 20 -- with ~140 candidates per sample the work is data-parallel already; removed.)
#ifndef PT_PAIR_CAP
#define PT_PAIR_CAP 192
#endif
constexpr int PAIR_CAP = PT_PAIR_CAP;  // staged candidates per sample (128 or 192); slot numbers stay 8-bit

__device__ __forceinline__ uint32_t seg16_min(uint32_t v)
{
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 8));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return v;
}

// grid_select for a half-warp: slots sl, sl + 16, ... below `total` (0: the half sits out).
// `amb` is per lane; the caller combines it over the half.
template <int C>
__device__ __forceinline__ void pair_select(const PointF *cand, uint32_t total, double qx, double qy,
                                            double qz, double r2, int k, unsigned sl, bool upper,
                                            uint32_t &mine, bool &amb)
{
    uint32_t h[C];
#pragma unroll
    for (int u = 0; u < C; ++u) {
        const uint32_t j = sl + 16u * u;
        uint32_t key = GKEY_NONE;
        if (j < total) {
            double px, py, pz;
            int pidx;
            GridRec<PointF>::load(cand, j, px, py, pz, pidx);
            const double d = dist2_exact(qx, qy, qz, px, py, pz);
            if (d <= r2) key = (__float_as_uint(__double2float_rd(d)) & ~0xffu) | j;
        }
        h[u] = key;
    }
    gkey_sort<C>(h);
    mine = GKEY_NONE;
    uint32_t m = GKEY_NONE;
#pragma unroll 2
    for (int r = 0; r <= k; ++r) {
        // two independent full-warp REDUX.MIN, each over one half's heads
        const uint32_t mA = __reduce_min_sync(0xffffffffu, upper ? GKEY_NONE : h[0]);
        const uint32_t mB = __reduce_min_sync(0xffffffffu, upper ? h[0] : GKEY_NONE);
        m = upper ? mB : mA;
        if ((unsigned)r == sl) mine = m;
        if ((m & 15u) == sl) {             // the owner pops its head
#pragma unroll
            for (int u = 0; u + 1 < C; ++u) h[u] = h[u + 1];
            h[C - 1] = GKEY_NONE;
        }
    }
    uint32_t nxt = __shfl_down_sync(0xffffffffu, mine, 1);
    if (sl == 15u) nxt = m;                     // k == 16: the 17th key is only in m
    amb = sl < (unsigned)k && mine != GKEY_NONE && (mine >> 8) == (nxt >> 8);
    if (sl >= (unsigned)k) mine = GKEY_NONE;
}

// grid_select_exact for the halves with act == true (the other half keeps its `mine`).
__device__ __forceinline__ void pair_select_exact(const PointF *cand, bool act, uint32_t total, double qx,
                                                  double qy, double qz, double r2, int k, unsigned lane,
                                                  uint32_t &mine)
{
    const unsigned sl = lane & 15u, hbase = lane & 16u;
    uint32_t taken = 0;
    if (act) mine = GKEY_NONE;
#pragma unroll 1
    for (int r = 0; r < k; ++r) {
        double bd = INFINITY;
        int bi = IDX_NONE, bu = -1;
        if (act) {
#pragma unroll 1
            for (int u = 0; sl + 16u * u < total; ++u) {
                if ((taken >> u) & 1u) continue;
                double px, py, pz;
                int pidx;
                GridRec<PointF>::load(cand, sl + 16u * u, px, py, pz, pidx);
                const double d = dist2_exact(qx, qy, qz, px, py, pz);
                if (d <= r2 && (bu < 0 || key_less(d, pidx, bd, bi))) { bd = d; bi = pidx; bu = u; }
            }
        }
        bool in = bu >= 0;
        if (!__any_sync(0xffffffffu, in)) break;
        const uint32_t hi = in ? (uint32_t)__double2hiint(bd) : 0xffffffffu;
        const uint32_t mh = seg16_min(hi);
        in = in && hi == mh;
        const uint32_t lo = in ? (uint32_t)__double2loint(bd) : 0xffffffffu;
        const uint32_t ml = seg16_min(lo);
        in = in && lo == ml;
        const uint32_t ii = in ? (uint32_t)bi : 0xffffffffu;
        const uint32_t mi = seg16_min(ii);
        in = in && ii == mi;
        const uint32_t seg = (__ballot_sync(0xffffffffu, in) >> hbase) & 0xffffu;
        const int w = (int)hbase + (seg ? __ffs(seg) - 1 : 0);
        const uint32_t jw = __shfl_sync(0xffffffffu, sl + 16u * (uint32_t)(bu < 0 ? 0 : bu), w);
        if (seg && lane == (unsigned)w) taken |= 1u << bu;
        if (seg && act && (unsigned)r == sl) mine = jw;
    }
}

// Outputs of both halves at once: sub-lane r holds the r-th neighbour of sample s (em: this
// half's answer is final).  Same arithmetic, in the same order, as the epilogue of grid_sample.
__device__ __forceinline__ void pair_emit(const QueryParams &P, uint32_t s, bool em, uint32_t mine, double d,
                                          int li, unsigned lane, PointF *wcand)
{
    const unsigned sl = lane & 15u, hbase = lane & 16u;
    const int k = P.k;
    const bool has = em && mine != GKEY_NONE;
    const int gid = has ? (P.ids ? __ldg(P.ids + li) : li) : -1;
    const bool want_blend = P.rgba_out || P.normal_out;
    AttrRaw at{0.f, 0.f, 0.f, 0u};
    if (has && (want_blend || P.cand_out) && P.attrs) at = load_attr(P.attrs + li);
    if (em && sl < (unsigned)k) {
        const size_t o = (size_t)s * k + sl;
        if (P.idx_out) P.idx_out[o] = gid;
        if (P.d2_out) P.d2_out[o] = d;
        if (P.cand_out) store_cand(P.cand_out + o, d, gid, at);
    }
    if (!want_blend) return;
    uint8_t *ro = P.rgba_out ? P.rgba_out + 4 * (size_t)s : nullptr;
    float *no = P.normal_out ? P.normal_out + 3 * (size_t)s : nullptr;
    const int cnt = __popc((__ballot_sync(0xffffffffu, has) >> hbase) & 0xffffu);
    if (em && cnt == 0 && sl == 0u) store_empty_blend(ro, no);
    const bool bl = em && cnt > 0;
    if (!__any_sync(0xffffffffu, bl)) return;
    const int mode = __shfl_sync(0xffffffffu, d, (int)hbase) == 0.0 ? 1 : 0;
    const double w = has ? (mode ? (d == 0.0 ? 1.0 : 0.0) : __ddiv_rn(1.0, d)) : 0.0;
    double *tile = reinterpret_cast<double *>(wcand);
    __syncwarp();                               // every lane has read its winner's record
    tile[0 * 32 + lane] = w;
    tile[1 * 32 + lane] = __dmul_rn(w, (double)(at.rgba & 0xffu));
    tile[2 * 32 + lane] = __dmul_rn(w, (double)((at.rgba >> 8) & 0xffu));
    tile[3 * 32 + lane] = __dmul_rn(w, (double)((at.rgba >> 16) & 0xffu));
    tile[4 * 32 + lane] = __dmul_rn(w, (double)at.nx);
    tile[5 * 32 + lane] = __dmul_rn(w, (double)at.ny);
    tile[6 * 32 + lane] = __dmul_rn(w, (double)at.nz);
    __syncwarp();
    double acc = 0.0;
    if (bl && sl < 7u) {
        const double *row = tile + sl * 32u + hbase;
        int j = 0;
#pragma unroll 1
        for (; j + 4 <= cnt; j += 4) {
            acc = __dadd_rn(acc, row[j]);
            acc = __dadd_rn(acc, row[j + 1]);
            acc = __dadd_rn(acc, row[j + 2]);
            acc = __dadd_rn(acc, row[j + 3]);
        }
#pragma unroll 1
        for (; j < cnt; ++j) acc = __dadd_rn(acc, row[j]);
    }
    double s0 = __shfl_sync(0xffffffffu, acc, (int)hbase);
    // (the nearest neighbour's attributes are fetched by every lane: only one half may need them,
    // and a *_sync shuffle must not sit in a branch that half a warp takes)
    const uint32_t c0 = __shfl_sync(0xffffffffu, at.rgba, (int)hbase);
    const float n0x = __shfl_sync(0xffffffffu, at.nx, (int)hbase), n0y = __shfl_sync(0xffffffffu, at.ny, (int)hbase),
                n0z = __shfl_sync(0xffffffffu, at.nz, (int)hbase);
    if (bl && !(s0 > 0.0 && s0 < INFINITY)) {   // overflowed weights: nearest neighbour only
        const double v[7] = {1.0, (double)(c0 & 0xffu), (double)((c0 >> 8) & 0xffu), (double)((c0 >> 16) & 0xffu),
                             (double)n0x, (double)n0y, (double)n0z};
        acc = 0.0;
#pragma unroll
        for (int a = 0; a < 7; ++a) if (sl == (unsigned)a) acc = v[a];
        s0 = 1.0;
    }
    const double s4 = __shfl_sync(0xffffffffu, acc, (int)hbase + 4), s5 = __shfl_sync(0xffffffffu, acc, (int)hbase + 5),
                 s6 = __shfl_sync(0xffffffffu, acc, (int)hbase + 6);
    const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(s4, s4), __dmul_rn(s5, s5)), __dmul_rn(s6, s6)));
    const double den = sl < 4u ? s0 : len;
    const double quo = __ddiv_rn(acc, den);
    const int ci = min(max(__double2int_rz(quo), 0), 255);
    const int cr = __shfl_sync(0xffffffffu, ci, (int)hbase + 1), cg = __shfl_sync(0xffffffffu, ci, (int)hbase + 2),
              cb = __shfl_sync(0xffffffffu, ci, (int)hbase + 3);
    if (bl && sl == 0u && ro) *reinterpret_cast<uchar4 *>(ro) = make_uchar4((unsigned char)cr, (unsigned char)cg, (unsigned char)cb, 255);
    if (bl && sl >= 4u && sl < 7u && no)
        no[sl - 4u] = (len > 0.0 && len < INFINITY) ? __double2float_rn(quo) : 0.0f;
}

#ifndef PT_PAIR_MIN_BLOCKS
#define PT_PAIR_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(GRID_WARPS * 32, PT_PAIR_MIN_BLOCKS)
knn_grid_pair_kernel(const QueryParams P, uint32_t *ovf_count, uint32_t *ovf_list)
{
    extern __shared__ __align__(128) unsigned char g_smem[];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned sl = lane & 15u, hbase = lane & 16u;
    PointF *wcand = reinterpret_cast<PointF *>(g_smem) + (size_t)wib * (2 * PAIR_CAP);
    PointF *cand = wcand + (hbase ? PAIR_CAP : 0);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(g_smem + sizeof(PointF) * 2 * PAIR_CAP * GRID_WARPS);
    const uint32_t bar = smem_u32(bars + wib);
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    uint32_t phase = 0;
    const int k = P.k;
    const GridParams &G = P.grid;
    const uint32_t n_warps = gridDim.x * GRID_WARPS, n_pairs = (P.m + 1u) >> 1;
    const PointF *pts = reinterpret_cast<const PointF *>(P.pts);
    uint32_t p = blockIdx.x * GRID_WARPS + wib;
    if (p >= n_pairs) return;
    while (p < n_pairs) {
        // a pair whose second sample does not exist answers the first one twice and emits it once
        const bool valid = 2u * p + (lane >> 4) < P.m;
        const uint32_t s = min(2u * p + (lane >> 4), P.m - 1u);
        const double qx = __ldg(P.queries + 3 * (size_t)s), qy = __ldg(P.queries + 3 * (size_t)s + 1),
                     qz = __ldg(P.queries + 3 * (size_t)s + 2);
        if (lane == 0 && p + n_warps < n_pairs)      // the next pair's coordinates: towards L2 meanwhile
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.queries + 6 * (size_t)(p + n_warps)));
        const double r2 = P.r2_per_query ? __ldg(P.r2_per_query + s) : P.r2;
        // a negative / NaN bound admits nothing: the empty answer is final (the unused rows of the
        // sharded path's fixed-capacity blocks arrive like this)
        const bool dead = valid && !(r2 >= 0.0);
        const bool look = valid && !dead;
        // ---- the sample's cell and its distance to the faces of the 3^3 block (see grid_sample) ---
        const GridTable &T = G.tab[G.att_tab[0]];
        const int sh = 21 - T.level;
        const uint32_t ncell = 1u << T.level;
        const float S = (float)(1u << sh), above = 2.0f * S;
        const double qv[3] = {qx, qy, qz};
        uint32_t c[3];
        float gl = INFINITY;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            const double t = (qv[ax] - G.lo[ax]) * G.inv_cell21;
            const uint32_t c21 = min(__double2uint_rz(t), 2097151u);
            c[ax] = c21 >> sh;
            const double f = t - (double)(c[ax] << sh);
            const float dl = __fadd_rd(__double2float_rd(f), S);
            const float du = __fsub_rd(above, __double2float_ru(f));
            gl = fminf(gl, c[ax] > 1u ? dl : INFINITY);
            gl = fminf(gl, c[ax] + 2u < ncell ? du : INFINITY);
        }
        // (|x| + |y| + |z| bounds the largest coordinate from above)
        const double slack = G.slack + 1.8e-15 * (fabs(qx) + fabs(qy) + fabs(qz));
        const double g = (double)gl * G.cell21 - slack;
        const double g2 = g > 0.0 ? __dmul_rd(g, g) : 0.0;
        // ---- look-ups: sub-lane c takes cells c and c + 16 ---------------------------------------
        uint32_t st0 = 0, ct0 = 0, st1 = 0, ct1 = 0;
        bool dense = false;
        {
            const uint32_t n1 = sl + 16u;
            const uint32_t x0 = c[0] + sl % 3u - 1u, y0 = c[1] + (sl / 3u) % 3u - 1u, z0 = c[2] + sl / 9u - 1u;
            const uint32_t x1 = c[0] + n1 % 3u - 1u, y1 = c[1] + (n1 / 3u) % 3u - 1u, z1 = c[2] + n1 / 9u - 1u;
            const bool ok0 = look && x0 < ncell && y0 < ncell && z0 < ncell;
            const bool ok1 = look && n1 < 27u && x1 < ncell && y1 < ncell && z1 < ncell;
            // (issuing both first probes before examining either was measured: the extra live
            // registers cost more than the second round trip)
            if (ok0) grid_lookup(T.buckets, T.cap, x0, y0, z0, st0, ct0);
            if (ok1) grid_lookup(T.buckets, T.cap, x1, y1, z1, st1, ct1);
            if (ct0 == 0xffffffffu) { dense = true; ct0 = 0; }
            if (ct1 == 0xffffffffu) { dense = true; ct1 = 0; }
        }
        const uint32_t mysum = ct0 + ct1;
        uint32_t incl = mysum;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (sl >= (unsigned)o) incl += y;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, (int)(hbase | 15u));
        const bool dense_h = ((__ballot_sync(0xffffffffu, dense) >> hbase) & 0xffffu) != 0u;
        const bool go = look && !dense_h && total <= (uint32_t)PAIR_CAP && (total >= (uint32_t)k || r2 <= g2);
        const uint32_t tot = go ? total : 0u;
        const uint32_t tA = __shfl_sync(0xffffffffu, tot, 0), tB = __shfl_sync(0xffffffffu, tot, 16);
#ifdef PT_STATS
        if (sl == 0 && valid) { atomicAdd(&g_stats[12], 1ull); atomicAdd(&g_stats[13], (unsigned long long)tot); }
#endif
        // ---- stage both halves' runs; one mbarrier ------------------------------------------------
        if (tA + tB > 0u) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_expect(bar, (tA + tB) * (uint32_t)sizeof(PointF));
            __syncwarp();
            const uint32_t excl = incl - mysum;
            if (go && ct0) bulk_g2s(smem_u32(cand + excl), pts + st0, ct0 * (uint32_t)sizeof(PointF), bar);
            if (go && ct1) bulk_g2s(smem_u32(cand + excl + ct0), pts + st1, ct1 * (uint32_t)sizeof(PointF), bar);
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        // ---- select --------------------------------------------------------------------------------
        uint32_t mine = GKEY_NONE;
        bool amb = false;
        const uint32_t tmax = max(tA, tB);
        if (tmax == 0u) { /* nothing staged (unused rows, empty blocks): no key to select */ }
        else if (tmax <= 32u) pair_select<2>(cand, tot, qx, qy, qz, r2, k, sl, hbase != 0u, mine, amb);
        else if (tmax <= 64u) pair_select<4>(cand, tot, qx, qy, qz, r2, k, sl, hbase != 0u, mine, amb);
#if PT_PAIR_CAP > 128
        else if (tmax <= 128u) pair_select<8>(cand, tot, qx, qy, qz, r2, k, sl, hbase != 0u, mine, amb);
        else pair_select<12>(cand, tot, qx, qy, qz, r2, k, sl, hbase != 0u, mine, amb);
#else
        else pair_select<8>(cand, tot, qx, qy, qz, r2, k, sl, hbase != 0u, mine, amb);
#endif
        const uint32_t amb_all = __ballot_sync(0xffffffffu, amb);
        if (amb_all) {
            const bool act = ((amb_all >> hbase) & 0xffffu) != 0u;
#ifdef PT_STATS
            if (sl == 0 && act) atomicAdd(&g_stats[14], 1ull);
#endif
            pair_select_exact(cand, act, tot, qx, qy, qz, r2, k, lane, mine);
        }
        double d = INFINITY;
        int li = IDX_NONE;
        if (mine != GKEY_NONE) {
            double px, py, pz;
            GridRec<PointF>::load(cand, mine & 0xffu, px, py, pz, li);
            d = dist2_exact(qx, qy, qz, px, py, pz);
        }
        const double kth = __shfl_sync(0xffffffffu, d, (int)hbase + k - 1);   // +inf while the list is short
        const bool done = dead || (go && fmin(kth, r2) <= g2);
        pair_emit(P, s, done, mine, d, li, lane, wcand);
        // ---- halves this attempt could not finish: the whole warp, the whole schedule ----------------
        const uint32_t redo = __ballot_sync(0xffffffffu, valid && !done);
        if (redo & 1u) {
            const double fx = __shfl_sync(0xffffffffu, qx, 0), fy = __shfl_sync(0xffffffffu, qy, 0),
                         fz = __shfl_sync(0xffffffffu, qz, 0);
            grid_sample<PointF, true>(P, 2u * p, fx, fy, fz, lane, wcand, bar, phase, ovf_count, ovf_list);
        }
        if (redo & 0x10000u) {
            const double fx = __shfl_sync(0xffffffffu, qx, 16), fy = __shfl_sync(0xffffffffu, qy, 16),
                         fz = __shfl_sync(0xffffffffu, qz, 16);
            grid_sample<PointF, true>(P, 2u * p + 1u, fx, fy, fz, lane, wcand, bar, phase, ovf_count, ovf_list);
        }
        p += n_warps;
    }
}

static inline size_t grid_kernel_smem(size_t rec_bytes)
{
    return rec_bytes * GRID_CAP * GRID_WARPS + 8 * GRID_WARPS;
}

static int launch_grid_pair(const QueryParams &qp, int sm_count, uint32_t *count, uint32_t *list, cudaStream_t s)
{
    const size_t smem = sizeof(PointF) * 2 * PAIR_CAP * GRID_WARPS + 8 * GRID_WARPS;
    if (smem > 48 * 1024)
        PT_CUDA(cudaFuncSetAttribute(knn_grid_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knn_grid_pair_kernel, GRID_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const unsigned want = ((qp.m + 1u) / 2u + GRID_WARPS - 1) / GRID_WARPS;
    const unsigned resident = (unsigned)(sm_count * per_sm);
    knn_grid_pair_kernel<<<want < resident ? want : resident, GRID_WARPS * 32, smem, s>>>(qp, count, list);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

template <typename PT>
static int launch_grid(const QueryParams &qp, int sm_count, uint32_t *count, uint32_t *list, cudaStream_t s)
{
    const size_t smem = grid_kernel_smem(sizeof(PT));
    const bool tma = opt_grid_tma() != 0;
    // two samples per warp when the lists fit a half-warp, the first attempt is a 3^3 block and
    // its candidates are expected to fit the half-warp's staging area with ~35 % to spare (a
    // sample with more is redone by the whole warp: correct, but the first pass was wasted)
    if (std::is_same<PT, PointF>::value && tma && opt_grid_pair() != 0 && qp.k <= 16 && qp.grid.n_attempts > 0 &&
        qp.grid.att_rc[0] == 1 && (opt_grid_pair() == 2 || qp.grid.expect_cand <= 0.74f * (float)PAIR_CAP)) {
        note_grid_pair_used(1);
        return launch_grid_pair(qp, sm_count, count, list, s);
    }
    note_grid_pair_used(0);
    auto kern = tma ? knn_grid_kernel<PT, true> : knn_grid_kernel<PT, false>;
    if (smem > 48 * 1024)
        PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GRID_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const unsigned want = (qp.m + GRID_WARPS - 1) / GRID_WARPS;
    const unsigned resident = (unsigned)(sm_count * per_sm);
    kern<<<want < resident ? want : resident, GRID_WARPS * 32, smem, s>>>(qp, count, list);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

}  // namespace pt
