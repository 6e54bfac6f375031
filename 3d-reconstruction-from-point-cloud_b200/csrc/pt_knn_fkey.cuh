// pt_knn_fkey.cuh -- variant 4 ("fkey"): the thread kernel (pt_knn_thread.cuh: one thread per
// sample, best-first over the 8-wide box pyramid, state in shared-memory columns) with an
// 8-byte top-k entry instead of the 12-byte (fp64 d2, index) pair:
//
//     entry = (fp32 key = d2 rounded DOWN) << 32 | position of the point in the sorted cloud
//
// Rounding is monotone, so key(a) < key(b) implies d2(a) < d2(b): every heap decision taken on
// strictly different keys is exact.  The only ambiguous case is EQUAL keys (two distances that
// agree to 2^-23 relative, or true ties); it is detected -- a candidate whose key equals the
// root's, or a root whose key equals one of its children's (in a max-heap a second entry with
// the root's key always has an ancestor chain of that key, i.e. a tied child of the root) --
// and decided on the exact (d2, index) re-read from the sorted cloud.  At the end the k
// winners' exact d2 and index are re-read as well and sorted by (d2, index) as before, so
// every output is bit-identical to variant 2.
// Gains: one 64-bit shared-memory access and one integer compare per heap level instead of
// two accesses and a three-instruction key compare, and 96 bytes less state per sample
// (k = 16), i.e. 17 instead of 14 resident warps per SM.
#pragma once

namespace pt {

constexpr uint32_t FKEY_INF = 0x7f800000u;

__device__ __forceinline__ uint32_t e_key(unsigned long long e) { return (uint32_t)(e >> 32); }

// sift `e` down from `pos` in the max-heap (by key) column of n entries
__device__ __forceinline__ void fheap_sift(unsigned long long *he, int pos, int n, unsigned long long e)
{
    const uint32_t ek = e_key(e);
    for (;;) {
        int c = 2 * pos + 1;
        if (c >= n) break;
        unsigned long long x = he[c * T_THREADS];
        if (c + 1 < n) {
            const unsigned long long y = he[(c + 1) * T_THREADS];
            if (e_key(y) > e_key(x)) { x = y; ++c; }
        }
        if (ek >= e_key(x)) break;
        he[pos * T_THREADS] = x;
        pos = c;
    }
    he[pos * T_THREADS] = e;
}

#ifndef PT_FKEY_MIN_BLOCKS
#define PT_FKEY_MIN_BLOCKS 20
#endif
template <typename PT>
__global__ void __launch_bounds__(T_THREADS, PT_FKEY_MIN_BLOCKS)
knn_fkey_kernel(const QueryParams P, uint32_t *ovf_count, uint32_t *ovf_list)
{
    extern __shared__ __align__(16) unsigned char t_smem[];
    const int k = P.k;
    const unsigned tid = threadIdx.x;
    unsigned long long *he = reinterpret_cast<unsigned long long *>(t_smem) + tid;        // [k]
    unsigned long long *pe = he + k * T_THREADS;                                          // [TPD_CAP]
    uint32_t *pqk = reinterpret_cast<uint32_t *>(pe + TPD_CAP * T_THREADS - tid) + tid;   // [TPQ_CAP]
    uint32_t *pqw = pqk + TPQ_CAP * T_THREADS;                                            // [TPQ_CAP]

    const uint32_t q = blockIdx.x * T_THREADS + tid;
    bool done = q >= P.m || P.t_levels == 0;
    bool overflow = false;

    double qx = 0, qy = 0, qz = 0, r2 = 0;
    if (q < P.m) {
        qx = __ldg(P.queries + 3 * (size_t)q);
        qy = __ldg(P.queries + 3 * (size_t)q + 1);
        qz = __ldg(P.queries + 3 * (size_t)q + 2);
        r2 = P.r2_per_query ? __ldg(P.r2_per_query + q) : P.r2;
    }
    const float qdn[3] = {__double2float_rd(qx), __double2float_rd(qy), __double2float_rd(qz)};
    const float qup[3] = {__double2float_ru(qx), __double2float_ru(qy), __double2float_ru(qz)};
    const float bound_r = __double2float_ru(r2);
    float bound = bound_r;

    int hn = 0;                      // candidates held; the column is a max-heap once hn == k
    uint32_t root_key = FKEY_INF;    // key of the heap root (current k-th) -- meaningful once hn == k
    int pq_n = 0;
    uint32_t lost = 0xffffffffu;   // smallest key of a queue entry that had to be given up
    const int qcap = min(max(P.pq_cap, 2), TPQ_CAP);   // runtime cap <= layout (tests shrink it)

    // drop the queue entries that lie beyond the bound (it only shrinks, so they are dead) and
    // rebuild the heap in place
    auto pq_compact = [&]() {
        int live = 0;
        for (int e = 0; e < pq_n; ++e) {
            const uint32_t ek = pqk[e * T_THREADS], ew = pqw[e * T_THREADS];
            if (__uint_as_float(ek & ~0xfu) <= bound) {
                int i = live++;
                while (i > 0) {
                    int p = (i - 1) >> 1;
                    uint32_t pk = pqk[p * T_THREADS];
                    if (pk <= ek) break;
                    pqk[i * T_THREADS] = pk;
                    pqw[i * T_THREADS] = pqw[p * T_THREADS];
                    i = p;
                }
                pqk[i * T_THREADS] = ek;
                pqw[i * T_THREADS] = ew;
            }
        }
        pq_n = live;
    };
    auto pq_push = [&](uint32_t key, uint32_t word) {
        if (pq_n == qcap) {
            pq_compact();
            if (pq_n == qcap) {
                // still full of live entries: give up the least promising one (the largest key;
                // in a min-heap it is among the leaves).  Exactness is kept by remembering the
                // smallest key ever given up: if the final bound stays below it, no dropped
                // subtree could have held a neighbour; otherwise the sample takes the fallback.
                int mi = qcap / 2;
                uint32_t mk = pqk[mi * T_THREADS];
#pragma unroll 1
                for (int e = qcap / 2 + 1; e < qcap; ++e) {
                    const uint32_t ek = pqk[e * T_THREADS];
                    if (ek > mk) { mk = ek; mi = e; }
                }
                if (key >= mk) { lost = min(lost, key); return; }
                lost = min(lost, mk);
                int i = mi;
                while (i > 0) {
                    int p = (i - 1) >> 1;
                    uint32_t pk = pqk[p * T_THREADS];
                    if (pk <= key) break;
                    pqk[i * T_THREADS] = pk;
                    pqw[i * T_THREADS] = pqw[p * T_THREADS];
                    i = p;
                }
                pqk[i * T_THREADS] = key;
                pqw[i * T_THREADS] = word;
                return;
            }
        }
        int i = pq_n++;
        while (i > 0) {
            int p = (i - 1) >> 1;
            uint32_t pk = pqk[p * T_THREADS];
            if (pk <= key) break;
            pqk[i * T_THREADS] = pk;
            pqw[i * T_THREADS] = pqw[p * T_THREADS];
            i = p;
        }
        pqk[i * T_THREADS] = key;
        pqw[i * T_THREADS] = word;
    };
    auto pq_pop = [&](uint32_t &key, uint32_t &word) {
        key = pqk[0];
        word = pqw[0];
        const int n = --pq_n;
        if (n == 0) return;
        const uint32_t lk = pqk[n * T_THREADS], lw = pqw[n * T_THREADS];
        int i = 0;
        for (;;) {
            int c = 2 * i + 1;
            if (c >= n) break;
            uint32_t ck = pqk[c * T_THREADS];
            if (c + 1 < n) {
                uint32_t ck2 = pqk[(c + 1) * T_THREADS];
                if (ck2 < ck) { ck = ck2; ++c; }
            }
            if (ck >= lk) break;
            pqk[i * T_THREADS] = ck;
            pqw[i * T_THREADS] = pqw[c * T_THREADS];
            i = c;
        }
        pqk[i * T_THREADS] = lk;
        pqw[i * T_THREADS] = lw;
    };
    // exact (d2, index) of an entry, re-read from the sorted cloud (rare: only on equal keys)
    auto exact_of = [&](unsigned long long e, double &d, int &idx) {
        double px, py, pz;
        PointLoad<PT>::load(P.pts, (uint32_t)e, px, py, pz, idx);
        d = dist2_exact(qx, qy, qz, px, py, pz);
    };
    // after the root changed: its key, and the tie check that keeps every later decision exact.
    // A second entry with the root's key always shows up as a tied child of the root (its
    // ancestors all carry that key); then every entry with that key is compared on the exact
    // (d2, index) and the largest becomes the root -- entries with equal keys can swap places
    // without breaking the heap.
    auto root_changed = [&]() {
        root_key = e_key(he[0]);
        const bool t1 = k > 1 && e_key(he[1 * T_THREADS]) == root_key;
        const bool t2 = k > 2 && e_key(he[2 * T_THREADS]) == root_key;
        if (!(t1 || t2)) return;
        double rd;
        int ri;
        exact_of(he[0], rd, ri);
        for (int j = 1; j < k; ++j) {
            const unsigned long long ej = he[j * T_THREADS];
            if (e_key(ej) != root_key) continue;
            double cd;
            int ci;
            exact_of(ej, cd, ci);
            if (key_less(rd, ri, cd, ci)) {
                he[j * T_THREADS] = he[0];
                he[0] = ej;
                rd = cd;
                ri = ci;
            }
        }
    };

    bool cur_valid = !done;
    int cur_tl = P.t_levels;
    uint32_t cur_id = 0, cur_mask = 0xffu;

    for (;;) {
        int leaf = -1;
        while (!done && leaf < 0) {
            if (!cur_valid) {
                if (pq_n == 0) { done = true; break; }
                uint32_t key, word;
                pq_pop(key, word);
                if (__uint_as_float(key & ~0xfu) > bound) { done = true; break; }  // rest is farther
                cur_tl = (int)(key & 0xfu);
                cur_id = word & 0x7fffffu;
                cur_mask = word >> 23;
            }
            cur_valid = false;
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
            if (best_c < 0) continue;
            rem &= ~(1u << best_c);
            if (rem) pq_push((__float_as_uint(second) & ~0xfu) | (uint32_t)cur_tl, (rem << 23) | cur_id);
            const uint32_t child = cur_id * 8 + (uint32_t)best_c;
            if (cur_tl == 1) {
                leaf = (int)child;
#if PT_T_PREFETCH
                {
                    const char *lp = reinterpret_cast<const char *>(P.pts) + (size_t)child * LEAF * sizeof(PT);
#pragma unroll
                    for (int l = 0; l < (int)(LEAF * sizeof(PT) / 128); ++l)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(lp + 128 * l));
                }
#endif
            } else {
                const bool dive = hn < k || pq_n == 0 ||
                                  best <= __uint_as_float(pqk[0] & ~0xfu);
                if (dive) { cur_valid = true; cur_tl -= 1; cur_id = child; cur_mask = 0xffu; }
                else pq_push((__float_as_uint(best) & ~0xfu) | (uint32_t)(cur_tl - 1),
                             (0xffu << 23) | child);
            }
            if (overflow) { done = true; leaf = -1; }
        }
        if (__all_sync(0xffffffffu, done)) break;

        // ---- leaf phase: every lane that holds a leaf scans it, 8 points per chunk -------------
        const uint32_t base = (uint32_t)(leaf < 0 ? 0 : leaf) * LEAF;
#pragma unroll 1
        for (int chunk = 0; chunk < LEAF / PT_T_CHUNK; ++chunk) {
            int pend = 0;
            if (leaf >= 0) {
#pragma unroll
                for (int p = 0; p < PT_T_CHUNK; ++p) {
                    const uint32_t pi = base + chunk * PT_T_CHUNK + p;
                    double px, py, pz;
                    int pidx;
                    PointLoad<PT>::load(P.pts, pi, px, py, pz, pidx);
                    const double d = dist2_exact(qx, qy, qz, px, py, pz);
                    const uint32_t cf = __float_as_uint(__double2float_rd(d));
                    if (pi < P.n && d <= r2 && (hn < k || cf <= root_key)) {
                        pe[pend * T_THREADS] = ((unsigned long long)cf << 32) | pi;
                        ++pend;
                    }
                }
            }
            while (__any_sync(0xffffffffu, pend > 0)) {
                if (pend > 0) {
                    --pend;
                    const unsigned long long e = pe[pend * T_THREADS];
                    if (hn < k) {
                        he[hn * T_THREADS] = e;
                        if (++hn == k) {
                            for (int s = k / 2 - 1; s >= 0; --s) fheap_sift(he, s, k, he[s * T_THREADS]);
                            root_changed();
                        }
                    } else if (e_key(e) < root_key) {
                        fheap_sift(he, 0, k, e);
                        root_changed();
                    } else if (e_key(e) == root_key) {   // undecidable on fp32 keys: exact compare
                        double cd, rd;
                        int ci, ri;
                        exact_of(e, cd, ci);
                        exact_of(he[0], rd, ri);
                        if (key_less(cd, ci, rd, ri)) {
                            fheap_sift(he, 0, k, e);
                            root_changed();
                        }
                    }
                }
            }
            if (overflow) { done = true; leaf = -1; }
        }
        // smallest fp32 value that is certainly >= the k-th exact d2 (its key is rounded down)
        if (hn == k) {
            const bool first = bound == bound_r;
            bound = fminf(__uint_as_float(root_key + 1u), bound_r);
            // the siblings queued during the first dive were pushed with an infinite bound:
            // most of them are dead now, which keeps a small queue sufficient
            if (first && bound < bound_r) pq_compact();
        }
    }

    // a dropped queue entry matters only if its subtree could still reach inside the final bound
    if (lost != 0xffffffffu && __uint_as_float(lost & ~0xfu) <= bound) overflow = true;
#ifdef PT_STATS
    {
        unsigned v = (q < P.m && overflow) ? 1u : 0u, w = q < P.m ? 1u : 0u;
        for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); w += __shfl_xor_sync(0xffffffffu, w, o); }
        if (tid == 0) { if (v) atomicAdd(&g_stats[8], (unsigned long long)v); atomicAdd(&g_stats[9], (unsigned long long)w); }
    }
#endif
    if (q >= P.m) return;
    if (overflow) {
        uint32_t slot = atomicAdd(ovf_count, 1u);
        ovf_list[slot] = q;
        return;
    }

    // the winners' exact (d2, index) from the sorted cloud, written over the dead entry /
    // pending / queue columns, then the common sort + output + blend
    double *hd = reinterpret_cast<double *>(t_smem) + tid;                                  // [k]
    int *hi = reinterpret_cast<int *>(t_smem + sizeof(double) * (size_t)k * T_THREADS) + tid;  // [k]
    // hd[j] takes the place of entry j; the int column starts after entry k-1 (pending / queue
    // columns, all dead by now), so the conversion is in place
    for (int j = 0; j < hn; ++j) {
        const uint32_t pi = (uint32_t)he[j * T_THREADS];
        double px, py, pz;
        int pidx;
        PointLoad<PT>::load(P.pts, pi, px, py, pz, pidx);
        hd[j * T_THREADS] = dist2_exact(qx, qy, qz, px, py, pz);
        hi[j * T_THREADS] = pidx;
    }
    emit_sample<T_THREADS>(P, q, hd, hi, hn);
}

static inline size_t fkey_kernel_smem(int k)
{
    // entries [k] + pending [TPD_CAP] (8 B each) + queue [TPQ_CAP] (8 B), and at least the
    // epilogue's (d2, index) columns, 12 B * k
    size_t per = (size_t)(k + TPD_CAP) * 8 + (size_t)TPQ_CAP * 8;
    if (per < (size_t)k * 12) per = (size_t)k * 12;
    return (size_t)T_THREADS * per;
}

template <typename PT>
static int launch_fkey(const QueryParams &qp, uint32_t *count, uint32_t *list, cudaStream_t s)
{
    static bool attr_set[2] = {false, false};
    const int which = sizeof(PT) == 32;
    if (!attr_set[which]) {
        PT_CUDA(cudaFuncSetAttribute(knn_fkey_kernel<PT>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[which] = true;
    }
    unsigned blocks = (qp.m + T_THREADS - 1) / T_THREADS;
    knn_fkey_kernel<PT><<<blocks, T_THREADS, fkey_kernel_smem(qp.k) + (size_t)opt_smem_pad(), s>>>(qp, count, list);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

}  // namespace pt
