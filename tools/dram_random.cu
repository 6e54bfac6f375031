// Achievable DRAM throughput for the access pattern of the grid kernel: random, short, 16-byte
// aligned chunks out of a multi-GB array (no reuse, nothing L2-resident).
//   usage: dram_random [array MiB] [chunk bytes] [loads in flight per thread]
// Prints chunk rate, useful GB/s and, with 64-byte granularity assumed, the sector GB/s.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// each group of `lanes_per_chunk` lanes reads one chunk (16 bytes per lane); `unroll` independent chunks in flight
template <int UNROLL>
__global__ void gather_kernel(const float4 *base, uint64_t n_slots, int lanes_per_chunk, int iters, float *out)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t grp = tid / lanes_per_chunk, within = tid % lanes_per_chunk;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t r = ((uint64_t)mix(grp * 2654435761u + it * UNROLL + u) << 20) ^ mix(grp + 977u * (it * UNROLL + u));
            const uint64_t slot = r % (n_slots - lanes_per_chunk);
            v[u] = __ldg(base + slot + within);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char **argv)
{
    const size_t mib = argc > 1 ? atol(argv[1]) : 4096;
    const int chunk = argc > 2 ? atoi(argv[2]) : 160;
    const int unroll = argc > 3 ? atoi(argv[3]) : 4;
    const int lanes = chunk / 16;
    const uint64_t n_slots = mib * (1ull << 20) / 16;
    float4 *buf; float *out;
    cudaMalloc(&buf, n_slots * 16); cudaMalloc(&out, 4);
    cudaMemset(buf, 1, n_slots * 16);
    const int threads = 256, blocks = 148 * 8 * 4, iters = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (unroll == 1) gather_kernel<1><<<blocks, threads>>>(buf, n_slots, lanes, iters, out);
        else if (unroll == 2) gather_kernel<2><<<blocks, threads>>>(buf, n_slots, lanes, iters, out);
        else if (unroll == 8) gather_kernel<8><<<blocks, threads>>>(buf, n_slots, lanes, iters, out);
        else gather_kernel<4><<<blocks, threads>>>(buf, n_slots, lanes, iters, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double chunks = (double)blocks * threads / lanes * iters * (unroll == 1 || unroll == 2 || unroll == 8 ? unroll : 4);
        const double span = chunk + 48.0;        // expected 64-byte lines touched by a chunk at a random 16-byte offset: (chunk + 48) / 64
        printf("array %zu MiB chunk %d B unroll %d: %.3f ms, %.2f G chunks/s, useful %.0f GB/s, 64-byte lines %.0f GB/s\n",
               mib, chunk, unroll, ms, chunks / ms * 1e-6, chunks * chunk / ms * 1e-6, chunks * span / ms * 1e-6);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
