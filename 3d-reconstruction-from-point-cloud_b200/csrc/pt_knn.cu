// pt_knn.cu -- exact k-NN query + fused blend (replaces the per-corner
// `K_neighbor_search search(tree, q, K)` loop, /root/reference src/pointsTransfer.cpp:470-479,
// metric src/Distance.h:6-11, pruning bound src/Distance.h:27-57).
//
// Variant 0 ("warp"): one warp per sample.  The warp walks the 32-wide levels of the box
// pyramid nearest-child-first; each lane tests one child box (conservative fp32 bound) or
// evaluates one point of a 32-point leaf (exact fp64 metric).  The running top-k is a sorted
// list distributed one entry per lane, so k <= 32.
#include "pt_index.cuh"

namespace pt {

// ---- warp-distributed sorted list ---------------------------------------------------------
struct WarpList {
    double d;     // lane j holds the j-th best (ascending); +inf when empty
    int    i;     // local point index; IDX_NONE when empty
    double kd;    // replicated: current k-th entry (the acceptance threshold)
    int    ki;
};

__device__ __forceinline__ double shfl_f64(double v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}

// Offer one candidate per lane (pass = lane has a candidate that beats the k-th entry).
__device__ __forceinline__ void warp_list_offer(WarpList &L, int k, unsigned lane, bool pass,
                                                double d, int idx)
{
    unsigned m = __ballot_sync(0xffffffffu, pass);
    while (m) {
        int c = __ffs(m) - 1;
        m &= m - 1;
        double cd = shfl_f64(d, c);
        int ci = __shfl_sync(0xffffffffu, idx, c);
        if (!key_less(cd, ci, L.kd, L.ki)) continue;  // threshold moved since the ballot
        unsigned before = __ballot_sync(0xffffffffu, key_less(L.d, L.i, cd, ci));
        unsigned pos = __popc(before);
        double ud = __shfl_up_sync(0xffffffffu, L.d, 1);
        int ui = __shfl_up_sync(0xffffffffu, L.i, 1);
        if (lane > pos) { L.d = ud; L.i = ui; }
        else if (lane == pos) { L.d = cd; L.i = ci; }
        L.kd = shfl_f64(L.d, k - 1);
        L.ki = __shfl_sync(0xffffffffu, L.i, k - 1);
    }
}

// ---- fused blend (frozen definition, DESIGN.md "blend") -------
__device__ __forceinline__ double butterfly_sum(double v)
{
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, s));
    return v;
}

// Lane j holds neighbour j (has = valid), its d2 and attribute record.
__device__ __forceinline__ void blend_store(unsigned lane, bool has, double d2, pt_attr at,
                                            uint8_t *rgba_out, float *normal_out)
{
    unsigned any = __ballot_sync(0xffffffffu, has);
    if (any == 0) {
        if (lane < 4 && rgba_out) rgba_out[lane] = 0;
        if (lane < 3 && normal_out) normal_out[lane] = 0.0f;
        return;
    }
    double d0 = shfl_f64(d2, 0);
    bool exact = d0 == 0.0;
    double w = 0.0;
    if (has) w = exact ? (d2 == 0.0 ? 1.0 : 0.0) : __ddiv_rn(1.0, d2);
    double W = butterfly_sum(w);
    if (!(W > 0.0 && W < INFINITY)) {  // overflowed weights: nearest neighbour only
        w = (lane == 0) ? 1.0 : 0.0;
        W = butterfly_sum(w);
    }
    double cr = butterfly_sum(__dmul_rn(w, (double)at.r));
    double cg = butterfly_sum(__dmul_rn(w, (double)at.g));
    double cb = butterfly_sum(__dmul_rn(w, (double)at.b));
    double sx = butterfly_sum(__dmul_rn(w, (double)at.nx));
    double sy = butterfly_sum(__dmul_rn(w, (double)at.ny));
    double sz = butterfly_sum(__dmul_rn(w, (double)at.nz));
    if (lane == 0) {
        if (rgba_out) {
            int r = __double2int_rz(__ddiv_rn(cr, W));
            int g = __double2int_rz(__ddiv_rn(cg, W));
            int b = __double2int_rz(__ddiv_rn(cb, W));
            uchar4 o;
            o.x = (unsigned char)min(max(r, 0), 255);
            o.y = (unsigned char)min(max(g, 0), 255);
            o.z = (unsigned char)min(max(b, 0), 255);
            o.w = 255;
            *reinterpret_cast<uchar4 *>(rgba_out) = o;
        }
        if (normal_out) {
            double len = __dsqrt_rn(
                __dadd_rn(__dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy)), __dmul_rn(sz, sz)));
            if (len > 0.0 && len < INFINITY) {
                normal_out[0] = __double2float_rn(__ddiv_rn(sx, len));
                normal_out[1] = __double2float_rn(__ddiv_rn(sy, len));
                normal_out[2] = __double2float_rn(__ddiv_rn(sz, len));
            } else {
                normal_out[0] = normal_out[1] = normal_out[2] = 0.0f;
            }
        }
    }
}

__device__ __forceinline__ pt_attr zero_attr()
{
    pt_attr a;
    a.nx = a.ny = a.nz = 0.f;
    a.r = a.g = a.b = a.a = 0;
    return a;
}

__device__ __forceinline__ pt_attr load_attr(const pt_attr *p)
{
    int4 v = __ldg(reinterpret_cast<const int4 *>(p));
    pt_attr a;
    a.nx = __int_as_float(v.x); a.ny = __int_as_float(v.y); a.nz = __int_as_float(v.z);
    a.r = v.w & 0xff; a.g = (v.w >> 8) & 0xff; a.b = (v.w >> 16) & 0xff; a.a = (v.w >> 24) & 0xff;
    return a;
}

// Writes every requested output of one sample from the warp-distributed list.
__device__ __forceinline__ void emit_sample(const QueryParams &P, uint32_t q, unsigned lane,
                                            double d, int li)
{
    const int k = P.k;
    bool has = lane < (unsigned)k && li != IDX_NONE;
    int gid = has ? (P.ids ? __ldg(P.ids + li) : li) : -1;
    if (lane < (unsigned)k) {
        size_t o = (size_t)q * k + lane;
        if (P.idx_out) P.idx_out[o] = gid;
        if (P.d2_out) P.d2_out[o] = has ? d : INFINITY;
    }
    bool need_attr = (P.rgba_out || P.normal_out || P.cand_out) && P.attrs;
    pt_attr at = zero_attr();
    if (need_attr && has) at = load_attr(P.attrs + li);
    if (P.cand_out && lane < (unsigned)k) {
        pt_cand c;
        c.d2 = has ? d : INFINITY;
        c.id = gid;
        c.r = at.r; c.g = at.g; c.b = at.b; c.a = at.a;
        c.nx = at.nx; c.ny = at.ny; c.nz = at.nz;
        c.pad_ = 0;
        P.cand_out[(size_t)q * k + lane] = c;
    }
    if (P.rgba_out || P.normal_out)
        blend_store(lane, has, d, at, P.rgba_out ? P.rgba_out + 4 * (size_t)q : nullptr,
                    P.normal_out ? P.normal_out + 3 * (size_t)q : nullptr);
}

// ---- variant 0: warp per sample -----------------------------------------------------------
template <typename PT> struct PointLoad;
template <> struct PointLoad<PointF> {
    static __device__ __forceinline__ void load(const void *base, uint32_t i, double &x, double &y,
                                                double &z, int &idx)
    {
        float4 v = __ldg(reinterpret_cast<const float4 *>(base) + i);
        x = (double)v.x; y = (double)v.y; z = (double)v.z;
        idx = __float_as_int(v.w);
    }
};
template <> struct PointLoad<PointD> {
    static __device__ __forceinline__ void load(const void *base, uint32_t i, double &x, double &y,
                                                double &z, int &idx)
    {
        const double2 *p = reinterpret_cast<const double2 *>(base) + 2 * (size_t)i;
        double2 a = __ldg(p), b = __ldg(p + 1);
        x = a.x; y = a.y; z = b.x;
        idx = (int)(__double_as_longlong(b.y) & 0xffffffffll);
    }
};

constexpr int WARPS_PER_BLOCK = 8;

template <typename PT>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) knn_warp_kernel(const QueryParams P)
{
    __shared__ float s_lb[WARPS_PER_BLOCK][MAX_W_LEVELS][32];
    const unsigned lane = threadIdx.x & 31;
    const unsigned wib = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * WARPS_PER_BLOCK + wib;
    if (q >= P.m) return;
    const int k = P.k;

    const double qx = __ldg(P.queries + 3 * (size_t)q);
    const double qy = __ldg(P.queries + 3 * (size_t)q + 1);
    const double qz = __ldg(P.queries + 3 * (size_t)q + 2);
    const float qdn[3] = {__double2float_rd(qx), __double2float_rd(qy), __double2float_rd(qz)};
    const float qup[3] = {__double2float_ru(qx), __double2float_ru(qy), __double2float_ru(qz)};
    const double r2 = P.r2_per_query ? __ldg(P.r2_per_query + q) : P.r2;

    WarpList L;
    L.d = INFINITY; L.i = IDX_NONE; L.kd = INFINITY; L.ki = IDX_NONE;
    float bound = __double2float_ru(r2);   // prune a box iff its lower bound > bound

    // per-level traversal state: lane l keeps level l's group id and pending-children mask
    uint32_t st_group = 0, st_mask = 0;
    int lvl = P.w_levels - 1;

    auto enter = [&](int level, uint32_t group) {
        const int pl = level * WLOG;
        const uint32_t node = group * 32 + lane;
        float lb = INFINITY;
        bool valid = node < P.pyr.count[pl];
        if (valid) {
            const int4 *bp = reinterpret_cast<const int4 *>(P.pyr.level[pl] + node);
            int4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
            Box b;
            b.lox = __int_as_float(b0.x); b.loy = __int_as_float(b0.y);
            b.loz = __int_as_float(b0.z); b.hix = __int_as_float(b0.w);
            b.hiy = __int_as_float(b1.x); b.hiz = __int_as_float(b1.y);
            lb = box_lower_bound(qdn, qup, b);
        }
        s_lb[wib][level][lane] = lb;
        unsigned mask = __ballot_sync(0xffffffffu, valid && lb <= bound);
        if (lane == (unsigned)level) { st_group = group; st_mask = mask; }
        __syncwarp();
    };

    if (lvl >= 0) enter(lvl, 0);
    while (lvl < P.w_levels && lvl >= 0) {
        unsigned mask = __shfl_sync(0xffffffffu, st_mask, lvl);
        if (mask == 0) { ++lvl; continue; }
        // nearest pending child first
        float lb = s_lb[wib][lvl][lane];
        unsigned bits = ((mask >> lane) & 1u) ? __float_as_uint(lb) : 0xffffffffu;
        unsigned mn = __reduce_min_sync(0xffffffffu, bits);
        if (__uint_as_float(mn) > bound) {  // nearest pending child is already too far
            if (lane == (unsigned)lvl) st_mask = 0;
            ++lvl;
            continue;
        }
        int c = __ffs(__ballot_sync(0xffffffffu, bits == mn)) - 1;
        if (lane == (unsigned)lvl) st_mask = mask & ~(1u << c);
        uint32_t node = __shfl_sync(0xffffffffu, st_group, lvl) * 32 + c;
        if (lvl > 0) {
            --lvl;
            enter(lvl, node);
            continue;
        }
        // leaf scan: one point per lane, exact metric (src/Distance.h:6-11)
        uint32_t i = node * LEAF + lane;
        double px, py, pz;
        int pidx;
        PointLoad<PT>::load(P.pts, i, px, py, pz, pidx);
        double d = dist2_exact(qx, qy, qz, px, py, pz);
        bool pass = i < P.n && d <= r2 && key_less(d, pidx, L.kd, L.ki);
        warp_list_offer(L, k, lane, pass, d, pidx);
        bound = __double2float_ru(fmin(L.kd, r2));
    }
    emit_sample(P, q, lane, L.d, L.i);
}

__global__ void __launch_bounds__(256) empty_result_kernel(const QueryParams P)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)P.m * P.k;
    if (i < total) {
        if (P.idx_out) P.idx_out[i] = -1;
        if (P.d2_out) P.d2_out[i] = INFINITY;
        if (P.cand_out) {
            pt_cand c{};
            c.d2 = INFINITY; c.id = -1;
            P.cand_out[i] = c;
        }
    }
    if (i < P.m) {
        if (P.rgba_out) *reinterpret_cast<uchar4 *>(P.rgba_out + 4 * i) = make_uchar4(0, 0, 0, 0);
        if (P.normal_out) { P.normal_out[3 * i] = P.normal_out[3 * i + 1] = P.normal_out[3 * i + 2] = 0.f; }
    }
}

int launch_query(pt_index *ix, const QueryParams &qp, cudaStream_t s)
{
    if (qp.m == 0) return PT_OK;
    if (qp.k < 1 || qp.k > PT_MAX_K) return PT_ERR_UNSUPPORTED;
    if (ix->n == 0) {
        size_t total = (size_t)qp.m * qp.k;
        empty_result_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(qp);
        count_launch();
        PT_CUDA(cudaGetLastError());
        return PT_OK;
    }
    unsigned blocks = (qp.m + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    if (ix->coord_f64)
        knn_warp_kernel<PointD><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(qp);
    else
        knn_warp_kernel<PointF><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(qp);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

// ---- K5: merge per-slab candidate lists (multi-GPU exchange epilogue) -------------------------
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
merge_kernel(const pt_cand *lists, int n_lists, uint32_t m, int k, int32_t *idx_out,
             double *d2_out, uint8_t *rgba_out, float *normal_out, pt_cand *cand_out)
{
    const unsigned lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (q >= m) return;
    WarpList L;
    L.d = INFINITY; L.i = IDX_NONE; L.kd = INFINITY; L.ki = IDX_NONE;
    int src = -1;  // where this lane's entry came from: list * 32 + slot
    for (int l = 0; l < n_lists; ++l) {
        const pt_cand *c = lists + ((size_t)l * m + q) * k + lane;
        double d = INFINITY;
        int id = IDX_NONE;
        if (lane < (unsigned)k) {
            d = c->d2;
            id = c->id;
            if (id < 0) { id = IDX_NONE; d = INFINITY; }
        }
        bool pass = id != IDX_NONE && key_less(d, id, L.kd, L.ki);
        // same insertion as warp_list_offer, carrying the source slot along
        unsigned mk = __ballot_sync(0xffffffffu, pass);
        while (mk) {
            int cl = __ffs(mk) - 1;
            mk &= mk - 1;
            double cd = shfl_f64(d, cl);
            int ci = __shfl_sync(0xffffffffu, id, cl);
            if (!key_less(cd, ci, L.kd, L.ki)) continue;
            unsigned before = __ballot_sync(0xffffffffu, key_less(L.d, L.i, cd, ci));
            unsigned pos = __popc(before);
            double ud = __shfl_up_sync(0xffffffffu, L.d, 1);
            int ui = __shfl_up_sync(0xffffffffu, L.i, 1);
            int us = __shfl_up_sync(0xffffffffu, src, 1);
            if (lane > pos) { L.d = ud; L.i = ui; src = us; }
            else if (lane == pos) { L.d = cd; L.i = ci; src = l * 32 + cl; }
            L.kd = shfl_f64(L.d, k - 1);
            L.ki = __shfl_sync(0xffffffffu, L.i, k - 1);
        }
    }
    bool has = lane < (unsigned)k && L.i != IDX_NONE;
    pt_cand mine{};
    mine.d2 = INFINITY; mine.id = -1;
    if (has) mine = lists[((size_t)(src >> 5) * m + q) * k + (src & 31)];
    if (lane < (unsigned)k) {
        size_t o = (size_t)q * k + lane;
        if (idx_out) idx_out[o] = has ? mine.id : -1;
        if (d2_out) d2_out[o] = has ? mine.d2 : INFINITY;
        if (cand_out) cand_out[o] = mine;
    }
    if (rgba_out || normal_out) {
        pt_attr at = zero_attr();
        if (has) { at.nx = mine.nx; at.ny = mine.ny; at.nz = mine.nz; at.r = mine.r; at.g = mine.g; at.b = mine.b; at.a = mine.a; }
        blend_store(lane, has, L.d, at, rgba_out ? rgba_out + 4 * (size_t)q : nullptr,
                    normal_out ? normal_out + 3 * (size_t)q : nullptr);
    }
}

int launch_merge(const pt_cand *lists, int n_lists, uint32_t m, int k, int32_t *idx_out,
                 double *d2_out, uint8_t *rgba_out, float *normal_out, pt_cand *cand_out,
                 cudaStream_t s)
{
    if (m == 0) return PT_OK;
    if (k < 1 || k > PT_MAX_K || n_lists < 1) return PT_ERR_UNSUPPORTED;
    unsigned blocks = (m + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    merge_kernel<<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(lists, n_lists, m, k, idx_out, d2_out,
                                                         rgba_out, normal_out, cand_out);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

}  // namespace pt
