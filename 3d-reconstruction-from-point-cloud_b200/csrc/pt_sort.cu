// pt_sort.cu -- K2: hand-written LSD radix sort of (64-bit key, 32-bit index) pairs.
//
// Sorts the Morton/Hilbert keys of the cloud.  8-bit digits; one
// pass = three kernels over tiles of 2048 keys:
//   rs_hist_kernel    per-tile digit histogram (per-warp shared-memory histograms)
//                     -> counts[digit][tile]
//   rs_scan_kernel    per digit, exclusive prefix over the tiles (block scan + carry), digit totals
//   rs_scatter_kernel stable in-tile ranking with __match_any_sync (a warp walks its keys in
//                     RS_ROUNDS coalesced rounds; tile-local rank = digit start in the tile +
//                     preceding warps + preceding rounds + lower lanes with the same digit), the
//                     tile is staged in digit order in shared memory and written out as runs
// Bytes moved per pass: 8 (hist read) + 12 (scatter read) + 12 (scatter write) per pair.
// HBM-bound by design; cub::DeviceRadixSort (CCCL, one-sweep) is the number it is compared
// against in profiles/ (pt_set_option("sort", 0) selects the library call for that comparison).
#include "pt_index.cuh"

namespace pt {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef PT_RS_ROUNDS
#define PT_RS_ROUNDS 8
#endif
constexpr int RS_ROUNDS = PT_RS_ROUNDS;             // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;     // 2048 keys: 24 KiB staged in shared memory
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const unsigned long long *keys, uint32_t n, int shift, uint32_t num_tiles,
               uint32_t *counts)
{
    // one private histogram per warp (plain shared-memory atomics: lanes of a warp rarely share
    // a digit, and when the whole warp does, one lane adds for all)
    __shared__ uint32_t h[RS_WARPS][RS_RADIX];
    const unsigned tid = threadIdx.x, w = tid >> 5;
#pragma unroll
    for (int j = 0; j < RS_WARPS; ++j) h[j][tid] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * (uint32_t)RS_TILE;
    uint32_t *hw = h[w];
    if (base + RS_TILE <= n) {
        // full tile: 16-byte loads, two keys per thread and round (order is irrelevant here)
        const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(keys + base);
        ulonglong2 v[RS_ROUNDS / 2];
#pragma unroll
        for (int it = 0; it < RS_ROUNDS / 2; ++it) v[it] = __ldg(k2 + it * RS_THREADS + tid);
#pragma unroll
        for (int it = 0; it < RS_ROUNDS / 2; ++it) {
            const unsigned d0 = (unsigned)(v[it].x >> shift) & 0xffu, d1 = (unsigned)(v[it].y >> shift) & 0xffu;
            const unsigned first = __shfl_sync(0xffffffffu, d0, 0);
            if (__all_sync(0xffffffffu, d0 == first && d1 == first)) {
                if ((tid & 31) == 0) atomicAdd(&hw[first], 64u);
            } else {
                atomicAdd(&hw[d0], 1u);
                atomicAdd(&hw[d1], 1u);
            }
        }
    } else {
        for (int it = 0; it < RS_ROUNDS; ++it) {
            const uint32_t i = base + it * RS_THREADS + tid;
            if (i < n) atomicAdd(&hw[(unsigned)(keys[i] >> shift) & 0xffu], 1u);
        }
    }
    __syncthreads();
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < RS_WARPS; ++j) c += h[j][tid];
    counts[(size_t)tid * num_tiles + blockIdx.x] = c;
}

// One block per digit: exclusive prefix of counts[digit][0..num_tiles) in place, total out.
__global__ void __launch_bounds__(RS_THREADS)
rs_scan_kernel(uint32_t *counts, uint32_t num_tiles, uint32_t *totals)
{
    __shared__ uint32_t warp_sum[RS_WARPS];
    __shared__ uint32_t carry_s;
    const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint32_t *row = counts + (size_t)blockIdx.x * num_tiles;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < num_tiles; t0 += RS_THREADS) {
        const uint32_t t = t0 + tid;
        const uint32_t v = t < num_tiles ? row[t] : 0;
        uint32_t x = v;                                  // inclusive warp scan
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, s);
            if (lane >= (unsigned)s) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int j = 0; j < RS_WARPS; ++j) wbase += j < (int)w ? warp_sum[j] : 0;
        const uint32_t carry = carry_s;
        if (t < num_tiles) row[t] = carry + wbase + x - v;
        __syncthreads();
        if (tid == RS_THREADS - 1) carry_s = carry + wbase + x;
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(RS_RADIX) rs_base_kernel(const uint32_t *totals, uint32_t *bases)
{
    __shared__ uint32_t s[RS_RADIX];
    const unsigned tid = threadIdx.x;
    s[tid] = totals[tid];
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int d = 0; d < RS_RADIX; ++d) { uint32_t v = s[d]; s[d] = acc; acc += v; }
    }
    __syncthreads();
    bases[tid] = s[tid];
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const unsigned long long *keys_in, const uint32_t *vals_in,
                  unsigned long long *keys_out, uint32_t *vals_out, uint32_t n, int shift,
                  uint32_t num_tiles, const uint32_t *counts, const uint32_t *bases)
{
    __shared__ uint32_t whist[RS_WARPS][RS_RADIX];   // per-warp digit counts, then running tile-local offsets
    __shared__ uint32_t gofs[RS_RADIX];              // digit d: global position of tile-local rank 0 of d, minus that rank
    __shared__ uint32_t wsum[RS_WARPS];
    __shared__ unsigned long long skey[RS_TILE];     // the tile in digit order (stable)
    __shared__ uint32_t sval[RS_TILE];
    const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int j = tid; j < RS_WARPS * RS_RADIX; j += RS_THREADS) (&whist[0][0])[j] = 0;
    __syncthreads();
    // warp w owns keys [w*32*ROUNDS, (w+1)*32*ROUNDS) of the tile, ROUNDS coalesced rounds of 32
    const uint32_t tile0 = blockIdx.x * (uint32_t)RS_TILE;
    const uint32_t wbase = tile0 + w * (32 * RS_ROUNDS);
    unsigned long long key[RS_ROUNDS];
    uint32_t val[RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        key[r] = i < n ? keys_in[i] : ~0ull;
        val[r] = i < n ? vals_in[i] : 0u;
    }
    // 1. per-warp histogram
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        const unsigned d = valid ? (unsigned)((key[r] >> shift) & 0xffu) : 0x100u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == (unsigned)(__ffs(peers) - 1)) whist[w][d] += (uint32_t)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // 2. digit d = tid: tile-local start of the digit (exclusive scan of the tile's digit
    //    totals), per-warp running offsets inside the tile, and the global offset of the run
    {
        const unsigned d = tid;
        uint32_t tot = 0;
#pragma unroll
        for (int j = 0; j < RS_WARPS; ++j) tot += whist[j][d];
        uint32_t x = tot;                                  // inclusive scan over the 256 digits
#pragma unroll
        for (int s2 = 1; s2 < 32; s2 <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, s2);
            if (lane >= (unsigned)s2) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        uint32_t pre = 0;
#pragma unroll
        for (int j = 0; j < RS_WARPS; ++j) pre += j < (int)w ? wsum[j] : 0;
        const uint32_t local_start = pre + x - tot;
        gofs[d] = bases[d] + counts[(size_t)d * num_tiles + blockIdx.x] - local_start;
        uint32_t off = local_start;
#pragma unroll
        for (int j = 0; j < RS_WARPS; ++j) {
            const uint32_t c = whist[j][d];
            whist[j][d] = off;
            off += c;
        }
    }
    __syncthreads();
    // 3. stable tile-local ranks; stage the tile in digit order in shared memory
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        const unsigned d = valid ? (unsigned)((key[r] >> shift) & 0xffu) : 0x100u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid) {
            const uint32_t rank = whist[w][d] + (uint32_t)__popc(peers & lt);
            skey[rank] = key[r];
            sval[rank] = val[r];
        }
        __syncwarp();
        if (valid && lane == (unsigned)(__ffs(peers) - 1)) whist[w][d] += (uint32_t)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // 4. write out: consecutive threads hold consecutive ranks, i.e. runs of one digit that go
    //    to consecutive global positions -- coalesced stores instead of one sector per pair
    const uint32_t cnt = min((uint32_t)RS_TILE, n - tile0);
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t j = r * RS_THREADS + tid;
        if (j < cnt) {
            const unsigned long long kj = skey[j];
            const uint32_t pos = gofs[(unsigned)((kj >> shift) & 0xffu)] + j;
            keys_out[pos] = kj;
            vals_out[pos] = sval[j];
        }
    }
}

size_t radix_sort_workspace_bytes(uint32_t n)
{
    const uint32_t num_tiles = (n + RS_TILE - 1) / RS_TILE;
    return sizeof(uint32_t) * ((size_t)RS_RADIX * num_tiles + 2 * RS_RADIX);
}

// Sorts pairs by key bits [first_bit, end_bit) (LSD, 8 bits per pass: the curve keys only need
// their top bits ordered, pt_build.cu).  Ping-pongs between (keys, vals) and
// (keys_alt, vals_alt); on return *keys_out / *vals_out point at the buffers holding the result.
int radix_sort_pairs(unsigned long long *keys, unsigned long long *keys_alt, uint32_t *vals,
                     uint32_t *vals_alt, uint32_t n, int first_bit, int end_bit, void *workspace,
                     cudaStream_t s, unsigned long long **keys_out, uint32_t **vals_out)
{
    *keys_out = keys;
    *vals_out = vals;
    if (n == 0) return PT_OK;
    const uint32_t num_tiles = (n + RS_TILE - 1) / RS_TILE;
    uint32_t *counts = (uint32_t *)workspace;
    uint32_t *totals = counts + (size_t)RS_RADIX * num_tiles;
    uint32_t *bases = totals + RS_RADIX;
    unsigned long long *kin = keys, *kout = keys_alt;
    uint32_t *vin = vals, *vout = vals_alt;
    for (int shift = first_bit; shift < end_bit; shift += 8) {
        rs_hist_kernel<<<num_tiles, RS_THREADS, 0, s>>>(kin, n, shift, num_tiles, counts);
        rs_scan_kernel<<<RS_RADIX, RS_THREADS, 0, s>>>(counts, num_tiles, totals);
        rs_base_kernel<<<1, RS_RADIX, 0, s>>>(totals, bases);
        rs_scatter_kernel<<<num_tiles, RS_THREADS, 0, s>>>(kin, vin, kout, vout, n, shift, num_tiles,
                                                         counts, bases);
        count_launch(4);
        unsigned long long *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    PT_CUDA(cudaGetLastError());
    *keys_out = kin;
    *vals_out = vin;
    return PT_OK;
}

}  // namespace pt
