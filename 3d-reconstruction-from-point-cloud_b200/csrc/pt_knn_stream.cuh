// pt_knn_stream.cuh -- variant 3 ("stream"): persistent one-warp blocks, one sample per lane,
// every lane an independent state machine that is advanced in warp-uniform ROUNDS:
//
//   refill : idle lanes take the next samples from a global counter (one atomic per warp), so
//            a lane never waits for the slowest sample of its warp and the kernel has no tail
//            of half-empty warps;
//   E step : one node expansion (8 child boxes) for every lane that still needs its NEXT leaf.
//            The traversal runs one leaf ahead of the scan with the bound it knows -- a stale
//            bound is only larger, so nothing is pruned wrongly; the leaf is re-checked against
//            the current bound before it is scanned;
//   S step : 8 points of the lane's current leaf, exact fp64 metric (src/Distance.h:6-11);
//            candidates that beat the k-th are parked;
//   D step : at most PT_S_INS parked candidates per lane go into the heap (root replacement +
//            sift-down; the heap is pre-filled with +inf so filling and replacing are one code
//            path);
//   flush  : a finished lane's heap column is copied to a global scratch row by the whole warp
//            (coalesced) and the lane becomes idle.
// Sorting, id mapping, the attribute gather and the blend run in knn_finalize_kernel (one
// thread per sample).  Same exactness core as variant 2: best-first order over the 8-wide box
// pyramid, fp32 bounds rounded toward the safe side, (d2, index) keys.
#pragma once

namespace pt {

#ifndef PT_S_INS
#define PT_S_INS 4
#endif
#ifndef PT_S_AHEAD
#define PT_S_AHEAD 2                // leaves the traversal may run ahead of the scan (1 or 2)
#endif
#ifndef PT_S_HEAP4
#define PT_S_HEAP4 1
#endif
#ifndef PT_S_REFILL_MASK
#define PT_S_REFILL_MASK 0          // refill when (round & mask) == 0
#endif
constexpr int SPQ_CAP = PT_TPQ_CAP;
constexpr int SPD_CAP = 8;
constexpr int S_CHUNK = 8;
constexpr int MARK_OVERFLOW = 0x7ffffffe;   // scratch row of a sample that goes to the fallback
constexpr int F_THREADS = 64;

template <int STRIDE>
__device__ __forceinline__ void heap_sift_s(double *hd, int *hi, int pos, int n, double cd, int ci)
{
    for (;;) {
        int c = 2 * pos + 1;
        if (c >= n) break;
        double xd = hd[c * STRIDE];
        int xi = hi[c * STRIDE];
        if (c + 1 < n) {
            double yd = hd[(c + 1) * STRIDE];
            int yi = hi[(c + 1) * STRIDE];
            if (key_less(xd, xi, yd, yi)) { xd = yd; xi = yi; ++c; }
        }
        if (!key_less(cd, ci, xd, xi)) break;
        hd[pos * STRIDE] = xd;
        hi[pos * STRIDE] = xi;
        pos = c;
    }
    hd[pos * STRIDE] = cd;
    hi[pos * STRIDE] = ci;
}

// Root replacement in a 4-ary max-heap of n entries (children of p: 4p+1 .. 4p+4): half the
// depth of the binary heap and the four child loads of a level are independent.
template <int STRIDE>
__device__ __forceinline__ void heap4_replace_root(double *hd, int *hi, int n, double cd, int ci)
{
    int pos = 0;
    for (;;) {
        const int c0 = 4 * pos + 1;
        if (c0 >= n) break;
        double bd = hd[c0 * STRIDE];
        int bi = hi[c0 * STRIDE];
        int bc = c0;
#pragma unroll
        for (int t = 1; t < 4; ++t) {
            const int c = c0 + t;
            if (c < n) {
                const double d = hd[c * STRIDE];
                const int i = hi[c * STRIDE];
                if (key_less(bd, bi, d, i)) { bd = d; bi = i; bc = c; }
            }
        }
        if (!key_less(cd, ci, bd, bi)) break;
        hd[pos * STRIDE] = bd;
        hi[pos * STRIDE] = bi;
        pos = bc;
    }
    hd[pos * STRIDE] = cd;
    hi[pos * STRIDE] = ci;
}

// Sorts the hn candidates of a column ascending by (d2, index) and writes every output of the
// sample: neighbour ids, d2, candidate records, blended colour / normal (frozen definition).
template <int STRIDE>
__device__ __forceinline__ void emit_sample(const QueryParams &P, uint32_t q, double *hd, int *hi,
                                            int hn)
{
    const int k = P.k;
    for (int s = hn / 2 - 1; s >= 0; --s)
        heap_sift_s<STRIDE>(hd, hi, s, hn, hd[s * STRIDE], hi[s * STRIDE]);
    for (int n = hn - 1; n > 0; --n) {
        const double ld = hd[n * STRIDE];
        const int li = hi[n * STRIDE];
        hd[n * STRIDE] = hd[0];
        hi[n * STRIDE] = hi[0];
        heap_sift_s<STRIDE>(hd, hi, 0, n, ld, li);
    }
    const bool want_blend = P.rgba_out || P.normal_out;
    const bool need_attr = (want_blend || P.cand_out) && P.attrs;
    const size_t o = (size_t)q * k;
    const int mode = (hn > 0 && hd[0] == 0.0) ? 1 : 0;
    BlendAcc acc;
    acc.reset();
    for (int j = 0; j < k; ++j) {
        const bool has = j < hn;
        const double d = has ? hd[j * STRIDE] : INFINITY;
        const int li = has ? hi[j * STRIDE] : IDX_NONE;
        const int gid = has ? (P.ids ? __ldg(P.ids + li) : li) : -1;
        if (P.idx_out) P.idx_out[o + j] = gid;
        if (P.d2_out) P.d2_out[o + j] = d;
        AttrRaw at{0.f, 0.f, 0.f, 0u};
        if (has && need_attr) at = load_attr(P.attrs + li);
        if (P.cand_out) store_cand(P.cand_out + o + j, d, gid, at);
        if (has && want_blend) acc.add(blend_weight(mode, d, j), at.rgba, at.nx, at.ny, at.nz);
    }
    if (want_blend) {
        uint8_t *ro = P.rgba_out ? P.rgba_out + 4 * (size_t)q : nullptr;
        float *no = P.normal_out ? P.normal_out + 3 * (size_t)q : nullptr;
        if (hn == 0) { store_empty_blend(ro, no); return; }
        if (!acc.weight_ok()) {   // overflowed weights: nearest neighbour only
            acc.reset();
            AttrRaw at = load_attr(P.attrs + hi[0]);
            acc.add(1.0, at.rgba, at.nx, at.ny, at.nz);
        }
        acc.store(ro, no);
    }
}

template <typename PT>
__global__ void __launch_bounds__(32)
knn_stream_kernel(const QueryParams P, uint32_t *ctl /* [0] overflow count, [1] next sample */,
                  uint32_t *ovf_list, double *scr_d, int *scr_i)
{
    extern __shared__ __align__(16) unsigned char s_smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int k = P.k;
    const unsigned lane = threadIdx.x;
    double *hd_base = reinterpret_cast<double *>(s_smem);                               // [k][32]
    double *pdd = hd_base + (size_t)k * 32 + lane;                                       // [SPD_CAP]
    int *hi_base = reinterpret_cast<int *>(s_smem + sizeof(double) * (k + SPD_CAP) * 32);
    int *pdi = hi_base + (size_t)k * 32 + lane;                                          // [SPD_CAP]
    uint32_t *pqk = reinterpret_cast<uint32_t *>(hi_base + (size_t)(k + SPD_CAP) * 32) + lane;
    uint32_t *pqw = pqk + SPQ_CAP * 32;                                                  // [SPQ_CAP]
    double *hd = hd_base + lane;
    int *hi = hi_base + lane;

#ifdef PT_STATS
    unsigned st_[16];
    for (int a = 0; a < 16; ++a) st_[a] = 0;
#endif

    // ---- per-lane sample state -----------------------------------------------------------
    int sid = -1;                 // sample being answered, -1 = idle
    bool pool_empty = false;      // warp-uniform
    double qx = 0, qy = 0, qz = 0, r2 = 0;
    float bound = 0.f;
    double root_d = INFINITY;     // heap root = current k-th (+inf until k candidates are held)
    int root_i = IDX_NONE;
    int pq_n = 0, pend = 0;
    bool cur_valid = false, trav_done = true, overflow = false;
    int cur_tl = 0;
    uint32_t cur_id = 0, cur_mask = 0;
    int scan_leaf = -1, scan_chunk = 0;
    int rdy_n = 0, rdy_leaf0 = 0, rdy_leaf1 = 0;      // leaves found ahead of the scan (FIFO)
    float rdy_lb0 = 0.f, rdy_lb1 = 0.f;

    for (int j = 0; j < k; ++j) { hd[j * 32] = INFINITY; hi[j * 32] = IDX_NONE; }
    unsigned req_mask = FULL;     // lanes waiting for the sample ids requested at the last flush
    uint32_t req_base = 0;        // lane 0: result of that request's atomicAdd
    if (lane == 0) req_base = atomicAdd(&ctl[1], 32u);

    auto pq_push = [&](uint32_t key, uint32_t word) {
        PT_STAT(2, 1);
        if (pq_n == SPQ_CAP) {
            // full: entries above the bound are dead -- drop them and rebuild; only a queue full
            // of live entries is an overflow (the sample then goes to the warp kernel)
            PT_STAT(4, 1);
            int live = 0;
            for (int e = 0; e < SPQ_CAP; ++e) {
                const uint32_t ek = pqk[e * 32], ew = pqw[e * 32];
                if (__uint_as_float(ek & ~0xfu) <= bound) {
                    int i = live++;
                    while (i > 0) {
                        int p = (i - 1) >> 1;
                        uint32_t pk = pqk[p * 32];
                        if (pk <= ek) break;
                        pqk[i * 32] = pk;
                        pqw[i * 32] = pqw[p * 32];
                        i = p;
                    }
                    pqk[i * 32] = ek;
                    pqw[i * 32] = ew;
                }
            }
            pq_n = live;
            if (pq_n == SPQ_CAP) { overflow = true; return; }
        }
        int i = pq_n++;
        while (i > 0) {
            int p = (i - 1) >> 1;
            uint32_t pk = pqk[p * 32];
            if (pk <= key) break;
            pqk[i * 32] = pk;
            pqw[i * 32] = pqw[p * 32];
            i = p;
        }
        pqk[i * 32] = key;
        pqw[i * 32] = word;
    };
    auto pq_pop = [&](uint32_t &key, uint32_t &word) {
        key = pqk[0];
        word = pqw[0];
        const int n = --pq_n;
        if (n == 0) return;
        const uint32_t lk = pqk[n * 32], lw = pqw[n * 32];
        int i = 0;
        for (;;) {
            int c = 2 * i + 1;
            if (c >= n) break;
            uint32_t ck = pqk[c * 32];
            if (c + 1 < n) {
                uint32_t ck2 = pqk[(c + 1) * 32];
                if (ck2 < ck) { ck = ck2; ++c; }
            }
            if (ck >= lk) break;
            pqk[i * 32] = ck;
            pqw[i * 32] = pqw[c * 32];
            i = c;
        }
        pqk[i * 32] = lk;
        pqw[i * 32] = lw;
    };

    for (unsigned round = 0;; ++round) {
        // nothing requested and nothing in flight: the pool is empty and every lane is idle
        if (req_mask == 0 && __all_sync(FULL, sid < 0)) break;
        if (lane == 0) PT_STAT(7, 1);

        // ---- S step, part 1: issue the 8 point loads of the current chunk; they are consumed
        // after the E step, so their latency overlaps the expansion
        const bool want_s = sid >= 0 && scan_leaf >= 0 && pend == 0 && !overflow;
        typename PointLoad<PT>::Raw praw[S_CHUNK];
        const uint32_t s_base = (uint32_t)(scan_leaf < 0 ? 0 : scan_leaf) * LEAF + (uint32_t)scan_chunk * S_CHUNK;
        if (want_s) {
#pragma unroll
            for (int p = 0; p < S_CHUNK; ++p) praw[p] = PointLoad<PT>::load_raw(P.pts, s_base + p);
        }

        // ---- E step: one expansion toward the next leaf ------------------------------------
        const bool want_e = sid >= 0 && !trav_done && rdy_n < PT_S_AHEAD && !overflow;
        if (__any_sync(FULL, want_e)) {
            if (lane == 0) PT_STAT(11, 1);
            if (want_e) {
                bool have = cur_valid;
                if (!have) {
                    if (pq_n == 0) trav_done = true;
                    else {
                        uint32_t key, word;
                        pq_pop(key, word);
                        PT_STAT(3, 1);
                        if (__uint_as_float(key & ~0xfu) > bound) { trav_done = true; pq_n = 0; }  // rest is farther
                        else {
                            cur_tl = (int)(key & 0xfu);
                            cur_id = word & 0x7fffffu;
                            cur_mask = word >> 23;
                            have = true;
                        }
                    }
                }
                cur_valid = false;
                if (have) {
                    PT_STAT(0, 1);
                    const float qdn[3] = {__double2float_rd(qx), __double2float_rd(qy), __double2float_rd(qz)};
                    const float qup[3] = {__double2float_ru(qx), __double2float_ru(qy), __double2float_ru(qz)};
                    const int pl = (cur_tl - 1) * T_LOG;
                    const uint32_t cnt = P.pyr.count[pl];
                    const Box *boxes = P.pyr.level[pl];
                    float best = INFINITY, second = INFINITY;
                    int best_c = -1;
                    uint32_t rem = 0;
                    Box cb[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) cb[c] = load_box(boxes + min(cur_id * 8 + c, cnt - 1));
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (((cur_mask >> c) & 1u) && cur_id * 8 + c < cnt) {
                            const float lb = box_lower_bound(qdn, qup, cb[c]);
                            if (lb <= bound) {
                                rem |= 1u << c;
                                if (lb < best) { second = best; best = lb; best_c = c; }
                                else second = fminf(second, lb);
                            }
                        }
                    }
                    if (best_c >= 0) {
                        rem &= ~(1u << best_c);
                        if (rem) pq_push((__float_as_uint(second) & ~0xfu) | (uint32_t)cur_tl, (rem << 23) | cur_id);
                        const uint32_t child = cur_id * 8 + (uint32_t)best_c;
                        if (cur_tl == 1) {
                            if (rdy_n == 0) { rdy_leaf0 = (int)child; rdy_lb0 = best; }
                            else { rdy_leaf1 = (int)child; rdy_lb1 = best; }
                            ++rdy_n;
                            const char *lp = reinterpret_cast<const char *>(P.pts) + (size_t)child * LEAF * sizeof(PT);
#pragma unroll
                            for (int l = 0; l < (int)(LEAF * sizeof(PT) / 128); ++l)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(lp + 128 * l));
                        } else {
                            const bool dive = root_d == INFINITY || pq_n == 0 ||
                                              best <= __uint_as_float(pqk[0] & ~0xfu);
                            if (dive) { cur_valid = true; cur_tl -= 1; cur_id = child; cur_mask = 0xffu; }
                            else pq_push((__float_as_uint(best) & ~0xfu) | (uint32_t)(cur_tl - 1),
                                         (0xffu << 23) | child);
                        }
                    }
                }
            }
        }

        // ---- refill: lanes that finished in the previous round get the sample ids requested
        // then (the atomic has been in flight since); their query loads are issued here and
        // first used at the end of this round / in the next E step
        if (req_mask) {
            const uint32_t base = __shfl_sync(FULL, req_base, 0);
            if (base + (uint32_t)__popc(req_mask) >= P.m) pool_empty = true;
            if ((req_mask >> lane) & 1u) {
                const uint32_t s = base + (uint32_t)__popc(req_mask & ((1u << lane) - 1u));
                if (s < P.m) {
                    sid = (int)s;
                    qx = __ldg(P.queries + 3 * (size_t)s);
                    qy = __ldg(P.queries + 3 * (size_t)s + 1);
                    qz = __ldg(P.queries + 3 * (size_t)s + 2);
                    r2 = P.r2_per_query ? __ldg(P.r2_per_query + s) : P.r2;
                    root_d = INFINITY; root_i = IDX_NONE;
                    pq_n = 0; pend = 0;
                    cur_valid = true; trav_done = false; overflow = false;
                    cur_tl = P.t_levels; cur_id = 0; cur_mask = 0xffu;
                    scan_leaf = -1; rdy_n = 0; scan_chunk = 0;
                    PT_STAT(9, 1);
                }
            }
            req_mask = 0;
        }

        // ---- S step, part 2: exact metric on the 8 points, park the candidates ---------------
        if (__any_sync(FULL, want_s)) {
            if (want_s) {
#pragma unroll
                for (int p = 0; p < S_CHUNK; ++p) {
                    const uint32_t pi = s_base + p;
                    double px, py, pz;
                    int pidx;
                    PointLoad<PT>::decode(praw[p], px, py, pz, pidx);
                    const double d = dist2_exact(qx, qy, qz, px, py, pz);
                    if (pi < P.n && d <= r2 && key_less(d, pidx, root_d, root_i)) {
                        pdd[pend * 32] = d;
                        pdi[pend * 32] = pidx;
                        ++pend;
                        PT_STAT(6, 1);
                    }
                }
                if (++scan_chunk == LEAF / S_CHUNK) scan_leaf = -1;
            }
        }

        // ---- D step: parked candidates -> heap ---------------------------------------------
#pragma unroll 1
        for (int it = 0; it < PT_S_INS; ++it) {
            if (!__any_sync(FULL, pend > 0)) break;
            if (lane == 0) PT_STAT(10, 1);
            if (pend > 0) {
                --pend;
                const double d = pdd[pend * 32];
                const int pidx = pdi[pend * 32];
                if (key_less(d, pidx, root_d, root_i)) {
                    PT_STAT(5, 1);
#if PT_S_HEAP4
                    heap4_replace_root<32>(hd, hi, k, d, pidx);
#else
                    heap_sift_s<32>(hd, hi, 0, k, d, pidx);
#endif
                    root_d = hd[0];
                    root_i = hi[0];
                }
            }
        }
        bound = __double2float_ru(fmin(root_d, r2));

        // ---- hand the next found leaf to the scanner (re-checked against the current bound) --
        if (sid >= 0 && scan_leaf < 0 && rdy_n > 0) {
            if (rdy_lb0 <= bound) { scan_leaf = rdy_leaf0; scan_chunk = 0; PT_STAT(1, 1); }
            rdy_leaf0 = rdy_leaf1; rdy_lb0 = rdy_lb1;
            --rdy_n;
        }

        // ---- flush finished lanes ----------------------------------------------------------
        const bool fin = sid >= 0 && (overflow || (trav_done && rdy_n == 0 && scan_leaf < 0 && pend == 0));
        unsigned fm = __ballot_sync(FULL, fin);
        if (fm) {
            __syncwarp();
            while (fm) {
                const int l = __ffs(fm) - 1;
                fm &= fm - 1;
                const int s_l = __shfl_sync(FULL, sid, l);
                const int ov_l = __shfl_sync(FULL, (int)overflow, l);
                if ((int)lane < k) {
                    const double d = hd_base[lane * 32 + l];
                    int i = hi_base[lane * 32 + l];
                    if (ov_l && lane == 0) i = MARK_OVERFLOW;
                    scr_d[(size_t)s_l * k + lane] = d;
                    scr_i[(size_t)s_l * k + lane] = i;
                    hd_base[lane * 32 + l] = INFINITY;
                    hi_base[lane * 32 + l] = IDX_NONE;
                }
            }
            __syncwarp();
            if (fin) {
                if (overflow) {
                    const uint32_t slot = atomicAdd(&ctl[0], 1u);
                    ovf_list[slot] = (uint32_t)sid;
                    PT_STAT(8, 1);
                }
                sid = -1;
            }
        }
        if (!pool_empty) {   // request sample ids for the lanes that are idle now (asynchronous)
            const unsigned idle_now = __ballot_sync(FULL, sid < 0);
            if (idle_now) {
                req_mask = idle_now;
                if (lane == 0) req_base = atomicAdd(&ctl[1], (uint32_t)__popc(idle_now));
            }
        }
    }
#ifdef PT_STATS
    for (int a = 0; a < 12; ++a) {
        unsigned v = st_[a];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0 && v) atomicAdd(&g_stats[a], (unsigned long long)v);
    }
#endif
}

// One thread per sample: scratch row -> sorted neighbours, ids, attributes, blend.
__global__ void __launch_bounds__(F_THREADS)
knn_finalize_kernel(const QueryParams P, const double *scr_d, const int *scr_i)
{
    extern __shared__ __align__(16) unsigned char f_smem[];
    const int k = P.k;
    const unsigned tid = threadIdx.x;
    double *hd = reinterpret_cast<double *>(f_smem) + tid;                                  // [k]
    int *hi = reinterpret_cast<int *>(f_smem + sizeof(double) * (size_t)k * F_THREADS) + tid;
    const uint32_t q = blockIdx.x * F_THREADS + tid;
    if (q >= P.m) return;
    const size_t o = (size_t)q * k;
    if (__ldg(scr_i + o) == MARK_OVERFLOW) return;     // answered by the fallback kernel
    int hn = 0;
    for (int j = 0; j < k; ++j) {
        const int i = __ldg(scr_i + o + j);
        const double d = __ldg(scr_d + o + j);
        if (i != IDX_NONE) {
            hd[hn * F_THREADS] = d;
            hi[hn * F_THREADS] = i;
            ++hn;
        }
    }
    emit_sample<F_THREADS>(P, q, hd, hi, hn);
}

static inline size_t stream_kernel_smem(int k)
{
    return (size_t)32 * ((size_t)(k + SPD_CAP) * 12 + (size_t)SPQ_CAP * 8);
}

template <typename PT>
static int launch_stream(const QueryParams &qp, uint32_t *ctl, uint32_t *list, double *scr_d,
                         int *scr_i, cudaStream_t s)
{
    static bool attr_set[2] = {false, false};
    static int sm_count = 0;
    const int which = sizeof(PT) == 32;
    if (!attr_set[which]) {
        PT_CUDA(cudaFuncSetAttribute(knn_stream_kernel<PT>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[which] = true;
    }
    if (!sm_count) {
        int dev = 0;
        PT_CUDA(cudaGetDevice(&dev));
        PT_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const size_t smem = stream_kernel_smem(qp.k) + (size_t)opt_smem_pad();
    int per_sm = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knn_stream_kernel<PT>, 32, smem));
    if (per_sm < 1) per_sm = 1;
    unsigned blocks = (qp.m + 31) / 32;
    const unsigned resident = (unsigned)(sm_count * per_sm);
    if (blocks > resident) blocks = resident;
    knn_stream_kernel<PT><<<blocks, 32, smem, s>>>(qp, ctl, list, scr_d, scr_i);
    count_launch();
    knn_finalize_kernel<<<(qp.m + F_THREADS - 1) / F_THREADS, F_THREADS,
                          (size_t)F_THREADS * 12 * qp.k, s>>>(qp, scr_d, scr_i);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

}  // namespace pt
