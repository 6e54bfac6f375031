// pt_knn_octet.cuh -- variant 1 ("octet", default): eight lanes per sample, four samples per warp.
//
// Included by pt_knn.cu after the shared helpers.  Mapping (B200: 32-wide warps, 148 SMs):
//   * box test   : the pyramid is walked 8-wide (levels 0,3,6,..); lane c of the octet tests
//                  child c of the popped node -- one coalesced 256-byte read of 8 boxes;
//   * leaf scan  : a 32-point leaf is one contiguous 512-byte (fp32) / 1-KiB (fp64) run; lane c
//                  reads points c, c+8, c+16, c+24 -- four fully coalesced 128-byte requests --
//                  and evaluates the exact fp64 metric (src/Distance.h:6-11) for its four;
//   * traversal  : best-first.  The per-sample priority queue of (fp32 box bound, node) lives
//                  in shared memory, unsorted; pop-min is an octet-wide reduction;
//   * top-k      : sorted list distributed over the octet, S = ceil(k/8) entries per lane in
//                  registers (blocked layout: lane l holds ranks l*S .. l*S+S-1).
// Samples whose queue overflows are appended to a list and re-run by the warp kernel.
#pragma once

namespace pt {

constexpr int O_THREADS = 128;
constexpr int O_PER_BLOCK = O_THREADS / 8;
constexpr int OPQ_CAP = 64;
constexpr int O_LOG = 3;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m)
{
    unsigned lo = __shfl_xor_sync(FULL, (unsigned)v, m);
    unsigned hi = __shfl_xor_sync(FULL, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}

template <int S>
struct OctetList {
    double d[S];
    int    i[S];
    double kd;   // octet-replicated: the k-th entry (acceptance threshold)
    int    ki;

    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int s = 0; s < S; ++s) { d[s] = INFINITY; i[s] = IDX_NONE; }
        kd = INFINITY; ki = IDX_NONE;
    }

    // All 32 lanes call this (it shuffles); `go` is octet-uniform: whether this octet inserts
    // candidate (cd, ci), which every lane of the octet holds.
    __device__ __forceinline__ void insert(bool go, unsigned l8, int k, double cd, int ci)
    {
        int c = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) c += key_less(d[s], i[s], cd, ci) ? 1 : 0;
        int pc = __shfl_up_sync(FULL, c, 1, 8);
        double od = __shfl_up_sync(FULL, d[S - 1], 1, 8);
        int oi = __shfl_up_sync(FULL, i[S - 1], 1, 8);
        if (l8 == 0) pc = S;
        if (go && c < S) {
            const bool ins = pc == S;     // the previous lane is entirely before the candidate
            const double nd = ins ? cd : od;
            const int ni = ins ? ci : oi;
#pragma unroll
            for (int s = S - 1; s > 0; --s)
                if (s > c) { d[s] = d[s - 1]; i[s] = i[s - 1]; }
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (s == c) { d[s] = nd; i[s] = ni; }
        }
        // refresh the threshold: rank k-1 lives in lane (k-1)/S, slot (k-1)%S
        const int ks = (k - 1) % S;
        double td = d[0];
        int ti = i[0];
#pragma unroll
        for (int s = 1; s < S; ++s)
            if (s == ks) { td = d[s]; ti = i[s]; }
        kd = __shfl_sync(FULL, td, (k - 1) / S, 8);
        ki = __shfl_sync(FULL, ti, (k - 1) / S, 8);
    }
};

template <typename PT, int S>
__global__ void __launch_bounds__(O_THREADS)
knn_octet_kernel(const QueryParams P, uint32_t *ovf_count, uint32_t *ovf_list)
{
    __shared__ uint32_t s_key[O_PER_BLOCK][OPQ_CAP];
    __shared__ uint32_t s_node[O_PER_BLOCK][OPQ_CAP];

    const unsigned lane = threadIdx.x & 31;
    const unsigned l8 = lane & 7;
    const unsigned oshift = lane & 24;             // first lane of this octet within the warp
    const unsigned oct_b = threadIdx.x >> 3;
    const int k = P.k;
    uint32_t *pk = s_key[oct_b];
    uint32_t *pn = s_node[oct_b];

    const uint32_t q = blockIdx.x * O_PER_BLOCK + oct_b;
    bool done = q >= P.m;
    bool overflow = false;

    double qx = 0, qy = 0, qz = 0, r2 = 0;
    if (!done) {
        qx = __ldg(P.queries + 3 * (size_t)q);
        qy = __ldg(P.queries + 3 * (size_t)q + 1);
        qz = __ldg(P.queries + 3 * (size_t)q + 2);
        r2 = P.r2_per_query ? __ldg(P.r2_per_query + q) : P.r2;
    }
    const float qdn[3] = {__double2float_rd(qx), __double2float_rd(qy), __double2float_rd(qz)};
    const float qup[3] = {__double2float_ru(qx), __double2float_ru(qy), __double2float_ru(qz)};
    float bound = __double2float_ru(r2);

    OctetList<S> L;
    L.init();
    int pq_n = 0;

    // lane c tests child c of node `id` (t-level tl; children at t-level tl-1) and the octet
    // appends the children that can still matter to its queue.  Warp-uniform call.
    auto expand = [&](bool go, int tl, uint32_t id) {
        float lb = 0.f;
        bool pass = false;
        const uint32_t cid = id * 8 + l8;
        if (go) {
            const int pl = (tl - 1) * O_LOG;
            if (cid < P.pyr.count[pl]) {
                lb = box_lower_bound(qdn, qup, load_box(P.pyr.level[pl] + cid));
                pass = lb <= bound;
            }
        }
        const unsigned b = (__ballot_sync(FULL, pass) >> oshift) & 0xffu;
        const int np = __popc(b);
        if (go && np) {
            if (pq_n + np > OPQ_CAP) overflow = true;
            else {
                if (pass) {
                    int pos = pq_n + __popc(b & ((1u << l8) - 1u));
                    // key = bound with its 4 low mantissa bits replaced by the child's level:
                    // still a valid (slightly smaller) lower bound, and among equal bounds the
                    // deeper node pops first, so zero-bound ties descend instead of fanning out
                    pk[pos] = (__float_as_uint(lb) & ~0xfu) | (uint32_t)(tl - 1);
                    pn[pos] = ((uint32_t)(tl - 1) << 28) | cid;
                }
                pq_n += np;
            }
        }
        __syncwarp();
    };

    expand(!done && P.t_levels > 0, P.t_levels, 0);

    for (;;) {
        // ---- pop the nearest pending node of every octet -------------------------------------
        unsigned long long best = ~0ull;
        if (!done && !overflow)
            for (int e = (int)l8; e < pq_n; e += 8)
                best = min(best, ((unsigned long long)pk[e] << 32) | (unsigned)e);
        best = min(best, shfl_xor_u64(best, 1));
        best = min(best, shfl_xor_u64(best, 2));
        best = min(best, shfl_xor_u64(best, 4));
        const uint32_t key = (uint32_t)(best >> 32);
        bool have = !done && !overflow && best != ~0ull &&
                    __uint_as_float(key & ~0xfu) <= bound;   // low 4 bits carry the level
        if (!have) done = true;
        if (__all_sync(FULL, done)) break;
        uint32_t node = 0;
        const int slot = (int)(best & 0xffffffffu);
        if (have) node = pn[slot];
        __syncwarp();
        if (have) {
            if (l8 == 0) { pk[slot] = pk[pq_n - 1]; pn[slot] = pn[pq_n - 1]; }
            --pq_n;
        }
        __syncwarp();
        const int tl = (int)(node >> 28);
        const uint32_t id = node & 0x0fffffffu;
        const bool is_int = have && tl > 0;
        const bool is_leaf = have && tl == 0;

        if (__any_sync(FULL, is_int)) expand(is_int, tl, id);

        if (__any_sync(FULL, is_leaf)) {
            // ---- leaf scan: 4 points per lane, exact metric ------------------------------------
            double cd[4];
            int ci[4];
            bool cp[4];
            const uint32_t base = id * LEAF;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cd[j] = INFINITY; ci[j] = IDX_NONE; cp[j] = false;
                if (is_leaf) {
                    const uint32_t pi = base + j * 8 + l8;
                    double px, py, pz;
                    PointLoad<PT>::load(P.pts, pi, px, py, pz, ci[j]);
                    cd[j] = dist2_exact(qx, qy, qz, px, py, pz);
                    cp[j] = pi < P.n && cd[j] <= r2 && key_less(cd[j], ci[j], L.kd, L.ki);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                unsigned m = (__ballot_sync(FULL, cp[j]) >> oshift) & 0xffu;
                while (__any_sync(FULL, m != 0)) {
                    const bool act = m != 0;
                    const int c = act ? __ffs(m) - 1 : 0;
                    m &= m - 1;
                    const double xd = __shfl_sync(FULL, cd[j], c, 8);
                    const int xi = __shfl_sync(FULL, ci[j], c, 8);
                    const bool go = act && key_less(xd, xi, L.kd, L.ki);
                    L.insert(go, l8, k, xd, xi);
                }
            }
            const float nb = __double2float_ru(fmin(L.kd, r2));
            const bool shrink = is_leaf && nb < bound;
            bound = nb;
            // The bound only ever decreases, so queue entries above it are dead: compact them
            // away (keeps the queue short for pop-min and far from its capacity).
            const int rounds = (__reduce_max_sync(FULL, shrink ? pq_n : 0) + 7) >> 3;
            int new_n = 0;
            for (int t = 0; t < rounds; ++t) {
                const int e = t * 8 + (int)l8;
                bool live = shrink && e < pq_n;
                uint32_t kk = 0, nn = 0;
                if (live) {
                    kk = pk[e];
                    nn = pn[e];
                    live = __uint_as_float(kk & ~0xfu) <= bound;
                }
                const unsigned b = (__ballot_sync(FULL, live) >> oshift) & 0xffu;
                __syncwarp();
                if (live) {
                    const int pos = new_n + __popc(b & ((1u << l8) - 1u));
                    pk[pos] = kk;
                    pn[pos] = nn;
                }
                new_n += __popc(b);
                __syncwarp();
            }
            if (shrink) pq_n = new_n;
        }
    }

    if (q >= P.m) return;
    if (overflow) {
        if (l8 == 0) {
            uint32_t at = atomicAdd(ovf_count, 1u);
            ovf_list[at] = q;
        }
        return;
    }

    // ---- outputs: rank e = l8*S + s -------------------------------------------------------------
    const bool want_blend = P.rgba_out || P.normal_out;
    const bool need_attr = (want_blend || P.cand_out) && P.attrs;
    const size_t o = (size_t)q * k;
    AttrRaw at[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int e = (int)l8 * S + s;
        const bool has = e < k && L.i[s] != IDX_NONE;
        at[s] = AttrRaw{0.f, 0.f, 0.f, 0u};
        if (e < k) {
            const int gid = has ? (P.ids ? __ldg(P.ids + L.i[s]) : L.i[s]) : -1;
            if (has && need_attr) at[s] = load_attr(P.attrs + L.i[s]);
            if (P.idx_out) P.idx_out[o + e] = gid;
            if (P.d2_out) P.d2_out[o + e] = has ? L.d[s] : INFINITY;
            if (P.cand_out) store_cand(P.cand_out + o + e, has ? L.d[s] : INFINITY, gid, at[s]);
        }
        if (!has) { L.d[s] = INFINITY; L.i[s] = IDX_NONE; }
    }
    if (!want_blend) return;
    // sequential blend in rank order; every lane of the octet accumulates the same sums.
    // NOTE: no early exit above for lanes of a valid sample, so the shuffles below are safe
    // within the octet (other octets of the warp may have returned: use the octet mask).
    const unsigned omask = 0xffu << oshift;
    const double d0 = __shfl_sync(omask, L.d[0], 0, 8);
    const int i0 = __shfl_sync(omask, L.i[0], 0, 8);
    uint8_t *ro = P.rgba_out ? P.rgba_out + 4 * (size_t)q : nullptr;
    float *no = P.normal_out ? P.normal_out + 3 * (size_t)q : nullptr;
    if (i0 == IDX_NONE) {
        if (l8 == 0) store_empty_blend(ro, no);
        return;
    }
    int mode = d0 == 0.0 ? 1 : 0;
    BlendAcc acc;
    for (int pass = 0; pass < 2; ++pass) {
        acc.reset();
        for (int l = 0; l < 8; ++l) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const double dj = __shfl_sync(omask, L.d[s], l, 8);
                const int ij = __shfl_sync(omask, L.i[s], l, 8);
                const uint32_t cj = __shfl_sync(omask, at[s].rgba, l, 8);
                const float nx = __shfl_sync(omask, at[s].nx, l, 8);
                const float ny = __shfl_sync(omask, at[s].ny, l, 8);
                const float nz = __shfl_sync(omask, at[s].nz, l, 8);
                const int e = l * S + s;
                if (e < k && ij != IDX_NONE) acc.add(blend_weight(mode, dj, e), cj, nx, ny, nz);
            }
        }
        if (acc.weight_ok()) break;
        mode = 2;
    }
    if (l8 == 0) acc.store(ro, no);
}

template <typename PT, int S>
static int launch_octet_s(const QueryParams &qp, uint32_t *count, uint32_t *list, cudaStream_t s)
{
    unsigned blocks = (qp.m + O_PER_BLOCK - 1) / O_PER_BLOCK;
    knn_octet_kernel<PT, S><<<blocks, O_THREADS, 0, s>>>(qp, count, list);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

template <typename PT>
static int launch_octet(const QueryParams &qp, uint32_t *count, uint32_t *list, cudaStream_t s)
{
    switch ((qp.k + 7) / 8) {
        case 1: return launch_octet_s<PT, 1>(qp, count, list, s);
        case 2: return launch_octet_s<PT, 2>(qp, count, list, s);
        case 3: return launch_octet_s<PT, 3>(qp, count, list, s);
        default: return launch_octet_s<PT, 4>(qp, count, list, s);
    }
}

}  // namespace pt
