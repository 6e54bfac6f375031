// pt_knn_persist.cuh -- variant 3 ("persistent thread"): the thread kernel's algorithm
// (pt_knn_thread.cuh) with persistent lanes.  In the one-sample-per-thread kernel a warp lasts
// as long as its slowest sample and finished lanes idle (13.9 of 32 lanes active on average,
// ncu).  Here a lane that finishes a sample dumps its k candidates (unsorted) to a scratch
// array and immediately takes the next sample of the warp's chunk; chunks are handed out by a
// global atomic counter.  A second, fully converged kernel (finalize) sorts every sample's
// candidates and produces the outputs + fused blend.
#pragma once

namespace pt {

#ifndef PT_PERSIST_CHUNK
#define PT_PERSIST_CHUNK 128
#endif

struct PersistScratch {
    double   *d;       // [k][m]  candidate squared distances (unsorted)
    int      *i;       // [k][m]  candidate local indices
    uint8_t  *n;       // [m]     candidates held; 255 = overflowed (re-run by the warp kernel)
    uint32_t *next;    // chunk counter
};

template <typename PT>
__global__ void __launch_bounds__(T_THREADS)
knn_persist_kernel(const QueryParams P, PersistScratch S, uint32_t *ovf_count, uint32_t *ovf_list)
{
    extern __shared__ __align__(16) unsigned char t_smem[];
    const int k = P.k;
    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31;
    double *hd = reinterpret_cast<double *>(t_smem) + tid;
    double *pdd = hd + k * T_THREADS;
    int *hi = reinterpret_cast<int *>(t_smem + sizeof(double) * (k + TPD_CAP) * T_THREADS) + tid;
    int *pdi = hi + k * T_THREADS;
    uint32_t *pqk = reinterpret_cast<uint32_t *>(pdi + TPD_CAP * T_THREADS);
    uint32_t *pqw = pqk + TPQ_CAP * T_THREADS;

    // per-sample state
    bool active = false;
    uint32_t q = 0;
    double qx = 0, qy = 0, qz = 0, r2 = 0;
    float qdn[3] = {0, 0, 0}, qup[3] = {0, 0, 0};
    float bound = 0.f;
    int hn = 0;
    double root_d = INFINITY;
    int root_i = IDX_NONE;
    int pq_n = 0;
    bool overflow = false;
    bool cur_valid = false;
    int cur_tl = 0;
    uint32_t cur_id = 0, cur_mask = 0;
    // warp-uniform chunk cursor
    uint32_t chunk_next = 0, chunk_end = 0;

    auto pq_push = [&](uint32_t key, uint32_t word) {
        if (pq_n == TPQ_CAP) {
            int live = 0;
            for (int e = 0; e < TPQ_CAP; ++e) {
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
            if (pq_n == TPQ_CAP) { overflow = true; return; }
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

    for (;;) {
        // ---- refill idle lanes from the warp's chunk (new chunk from the global counter) -------
        const unsigned need = __ballot_sync(0xffffffffu, !active);
        if (need) {
            if (chunk_next >= chunk_end) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(S.next, (uint32_t)PT_PERSIST_CHUNK);
                c = __shfl_sync(0xffffffffu, c, 0);
                chunk_next = min(c, P.m);
                chunk_end = min(c + (uint32_t)PT_PERSIST_CHUNK, P.m);
            }
            const uint32_t avail = chunk_end - chunk_next;
            const uint32_t rank = __popc(need & ((1u << lane) - 1u));
            if (!active && rank < avail) {
                q = chunk_next + rank;
                qx = __ldg(P.queries + 3 * (size_t)q);
                qy = __ldg(P.queries + 3 * (size_t)q + 1);
                qz = __ldg(P.queries + 3 * (size_t)q + 2);
                r2 = P.r2_per_query ? __ldg(P.r2_per_query + q) : P.r2;
                qdn[0] = __double2float_rd(qx); qdn[1] = __double2float_rd(qy); qdn[2] = __double2float_rd(qz);
                qup[0] = __double2float_ru(qx); qup[1] = __double2float_ru(qy); qup[2] = __double2float_ru(qz);
                bound = __double2float_ru(r2);
                hn = 0; root_d = INFINITY; root_i = IDX_NONE; pq_n = 0; overflow = false;
                cur_valid = true; cur_tl = P.t_levels; cur_id = 0; cur_mask = 0xffu;
                active = true;
            }
            chunk_next += min((uint32_t)__popc(need), avail);
        }
        if (!__any_sync(0xffffffffu, active)) break;

        // ---- traverse until this lane holds a leaf or its sample is finished ---------------------
        int leaf = -1;
        bool finished = false;
        while (active && !finished && leaf < 0) {
            if (!cur_valid) {
                if (pq_n == 0) { finished = true; break; }
                uint32_t key, word;
                pq_pop(key, word);
                if (__uint_as_float(key & ~0xfu) > bound) { finished = true; break; }
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
            } else {
                const bool dive = hn < k || pq_n == 0 ||
                                  best <= __uint_as_float(pqk[0] & ~0xfu);
                if (dive) { cur_valid = true; cur_tl -= 1; cur_id = child; cur_mask = 0xffu; }
                else pq_push((__float_as_uint(best) & ~0xfu) | (uint32_t)(cur_tl - 1),
                             (0xffu << 23) | child);
            }
            if (overflow) { finished = true; leaf = -1; }
        }

        // ---- leaf phase ----------------------------------------------------------------------------
        if (__any_sync(0xffffffffu, leaf >= 0)) {
            const uint32_t base = (uint32_t)(leaf < 0 ? 0 : leaf) * LEAF;
#pragma unroll 1
            for (int chunk = 0; chunk < LEAF / 8; ++chunk) {
                int pend = 0;
                if (leaf >= 0) {
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const uint32_t pi = base + chunk * 8 + p;
                        double px, py, pz;
                        int pidx;
                        PointLoad<PT>::load(P.pts, pi, px, py, pz, pidx);
                        const double d = dist2_exact(qx, qy, qz, px, py, pz);
                        if (pi < P.n && d <= r2 && (hn < k || key_less(d, pidx, root_d, root_i))) {
                            pdd[pend * T_THREADS] = d;
                            pdi[pend * T_THREADS] = pidx;
                            ++pend;
                        }
                    }
                }
                while (__any_sync(0xffffffffu, pend > 0)) {
                    if (pend > 0) {
                        --pend;
                        const double d = pdd[pend * T_THREADS];
                        const int pidx = pdi[pend * T_THREADS];
                        if (hn < k) {
                            hd[hn * T_THREADS] = d;
                            hi[hn * T_THREADS] = pidx;
                            if (++hn == k) {
                                heapify(hd, hi, k);
                                root_d = hd[0];
                                root_i = hi[0];
                            }
                        } else if (key_less(d, pidx, root_d, root_i)) {
                            heap_sift(hd, hi, 0, k, d, pidx);
                            root_d = hd[0];
                            root_i = hi[0];
                        }
                    }
                }
            }
            if (leaf >= 0 && hn == k) bound = __double2float_ru(fmin(root_d, r2));
        }

        // ---- finished lanes park their candidates and become idle ------------------------------------
        if (active && finished) {
            if (overflow) {
                S.n[q] = 255;
                ovf_list[atomicAdd(ovf_count, 1u)] = q;
            } else {
                for (int j = 0; j < hn; ++j) {
                    S.d[(size_t)j * P.m + q] = hd[j * T_THREADS];
                    S.i[(size_t)j * P.m + q] = hi[j * T_THREADS];
                }
                S.n[q] = (uint8_t)hn;
            }
            active = false;
        }
    }
}

// Sorts every sample's parked candidates (heap-sort in a shared-memory column) and writes the
// outputs + fused blend.  One thread per sample, fully converged.
template <int DUMMY = 0>
__global__ void __launch_bounds__(T_THREADS) knn_finalize_kernel(const QueryParams P, PersistScratch S)
{
    extern __shared__ __align__(16) unsigned char t_smem[];
    const int k = P.k;
    const unsigned tid = threadIdx.x;
    double *hd = reinterpret_cast<double *>(t_smem) + tid;
    int *hi = reinterpret_cast<int *>(t_smem + sizeof(double) * k * T_THREADS) + tid;
    const uint32_t q = blockIdx.x * T_THREADS + tid;
    if (q >= P.m) return;
    const int hn = S.n[q];
    if (hn == 255) return;     // overflowed: the warp kernel produces this sample
    for (int j = 0; j < hn; ++j) {
        hd[j * T_THREADS] = S.d[(size_t)j * P.m + q];
        hi[j * T_THREADS] = S.i[(size_t)j * P.m + q];
    }
    heapify(hd, hi, hn);
    for (int n = hn - 1; n > 0; --n) {
        const double ld = hd[n * T_THREADS];
        const int li = hi[n * T_THREADS];
        hd[n * T_THREADS] = hd[0];
        hi[n * T_THREADS] = hi[0];
        heap_sift(hd, hi, 0, n, ld, li);
    }
    const bool want_blend = P.rgba_out || P.normal_out;
    const bool need_attr = (want_blend || P.cand_out) && P.attrs;
    const size_t o = (size_t)q * k;
    const int mode = (hn > 0 && hd[0] == 0.0) ? 1 : 0;
    BlendAcc acc;
    acc.reset();
    for (int j = 0; j < k; ++j) {
        const bool has = j < hn;
        const double d = has ? hd[j * T_THREADS] : INFINITY;
        const int li = has ? hi[j * T_THREADS] : IDX_NONE;
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
        if (!acc.weight_ok()) {
            acc.reset();
            AttrRaw at = load_attr(P.attrs + hi[0]);
            acc.add(1.0, at.rgba, at.nx, at.ny, at.nz);
        }
        acc.store(ro, no);
    }
}

static inline size_t persist_scratch_bytes(uint32_t m, int k)
{
    size_t a = ((size_t)m * k * 8 + 255) & ~(size_t)255;
    size_t b = ((size_t)m * k * 4 + 255) & ~(size_t)255;
    size_t c = ((size_t)m + 255) & ~(size_t)255;
    return a + b + c + 256;
}

template <typename PT>
static int launch_persist(const QueryParams &qp, void *scratch, uint32_t *count, uint32_t *list,
                          cudaStream_t s)
{
    static bool attr_set[2] = {false, false};
    static int blocks_per_sm[2] = {0, 0};
    static int n_sm = 0;
    const int which = sizeof(PT) == 32;
    if (!attr_set[which]) {
        PT_CUDA(cudaFuncSetAttribute(knn_persist_kernel<PT>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)thread_kernel_smem(PT_MAX_K)));
        PT_CUDA(cudaFuncSetAttribute(knn_finalize_kernel<0>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     PT_MAX_K * T_THREADS * 12));
        int dev = 0;
        PT_CUDA(cudaGetDevice(&dev));
        PT_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        attr_set[which] = true;
    }
    PersistScratch S;
    char *p = (char *)scratch;
    S.d = (double *)p; p += ((size_t)qp.m * qp.k * 8 + 255) & ~(size_t)255;
    S.i = (int *)p;    p += ((size_t)qp.m * qp.k * 4 + 255) & ~(size_t)255;
    S.n = (uint8_t *)p; p += ((size_t)qp.m + 255) & ~(size_t)255;
    S.next = (uint32_t *)p;
    PT_CUDA(cudaMemsetAsync(S.next, 0, sizeof(uint32_t), s));
    const size_t smem = thread_kernel_smem(qp.k);
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[which],
                                                          knn_persist_kernel<PT>, T_THREADS, smem));
    unsigned need = (qp.m + T_THREADS - 1) / T_THREADS;
    unsigned blocks = (unsigned)(n_sm * (blocks_per_sm[which] > 0 ? blocks_per_sm[which] : 1));
    if (blocks > need) blocks = need;
    knn_persist_kernel<PT><<<blocks, T_THREADS, smem, s>>>(qp, S, count, list);
    count_launch();
    PT_CUDA(cudaGetLastError());
    knn_finalize_kernel<0><<<need, T_THREADS, (size_t)qp.k * T_THREADS * 12, s>>>(qp, S);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

}  // namespace pt
