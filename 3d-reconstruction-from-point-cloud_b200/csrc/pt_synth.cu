// pt_synth.cu -- synthetic clouds / mesh samples for tests and bench (include/pt_synth.h;
// SURVEY.md section 8 row M1).  Counter-based RNG: Philox4x32-10, key = seed, counter = global
// point index, so any slab can be regenerated independently on any rank.
//
// Bench / test scaffolding: built into its OWN library (libpt_synth_b200.so), not into the
// drop-in libpoints_transfer_b200.so, and its launches are not counted as product kernels.
#include <cmath>
#include <cstdint>
#include <cstdio>

#include <cuda_runtime.h>

#include "pt_synth.h"

#define PT_CUDA(call)                                                       \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) return e__ == cudaErrorNoDevice ? PT_ERR_NO_DEVICE : PT_ERR_CUDA; \
    } while (0)
static inline void count_launch() {}

namespace pt {

struct Philox {
    uint32_t c[4];
    uint32_t k[2];
};

__host__ __device__ inline void philox_round(Philox &s)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint64_t p0 = (uint64_t)M0 * s.c[0];
    uint64_t p1 = (uint64_t)M1 * s.c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ s.c[1] ^ s.k[0];
    uint32_t n2 = hi0 ^ s.c[3] ^ s.k[1];
    s.c[0] = n0; s.c[1] = lo1; s.c[2] = n2; s.c[3] = lo0;
}

__host__ __device__ inline void philox4x32_10(uint64_t counter, uint32_t stream, uint64_t seed,
                                              uint32_t out[4])
{
    Philox s;
    s.c[0] = (uint32_t)counter; s.c[1] = (uint32_t)(counter >> 32); s.c[2] = stream; s.c[3] = 0;
    s.k[0] = (uint32_t)seed; s.k[1] = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(s);
        s.k[0] += 0x9E3779B9u;
        s.k[1] += 0xBB67AE85u;
    }
    out[0] = s.c[0]; out[1] = s.c[1]; out[2] = s.c[2]; out[3] = s.c[3];
}

__host__ __device__ inline double u01(uint32_t a, uint32_t b)
{   // 53-bit uniform in [0,1)
    return (double)((((uint64_t)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}

// Heightfield z(x, y) = sum_i a_i sin(f_i x + p_i) sin(g_i y + q_i) and its gradient.
__host__ __device__ inline void surface(double x, double y, double &z, double &zx, double &zy)
{
    const double A[4] = {12.0, 6.0, 2.5, 0.8};
    const double F[4] = {0.021, 0.047, 0.11, 0.31};
    const double G[4] = {0.017, 0.039, 0.13, 0.27};
    const double Pp[4] = {0.3, 1.7, 2.9, 0.5};
    const double Qp[4] = {1.1, 0.2, 4.1, 3.3};
    z = 0; zx = 0; zy = 0;
    for (int i = 0; i < 4; ++i) {
        double sx = sin(F[i] * x + Pp[i]), cx = cos(F[i] * x + Pp[i]);
        double sy = sin(G[i] * y + Qp[i]), cy = cos(G[i] * y + Qp[i]);
        z += A[i] * sx * sy;
        zx += A[i] * F[i] * cx * sy;
        zy += A[i] * G[i] * sx * cy;
    }
}

__device__ inline void gauss2(uint32_t a, uint32_t b, uint32_t c, uint32_t d, double &g0, double &g1)
{
    double u1 = u01(a, b), u2 = u01(c, d);
    double r = sqrt(-2.0 * log(1.0 - u1));   // 1-u1 in (0,1]
    g0 = r * cos(6.283185307179586 * u2);
    g1 = r * sin(6.283185307179586 * u2);
}

__device__ inline pt_attr make_attr(double x, double y, double nx, double ny, double nz)
{
    pt_attr a;
    a.nx = (float)nx; a.ny = (float)ny; a.nz = (float)nz;
    a.r = (uint8_t)(127.5 + 127.4 * sin(0.05 * x));
    a.g = (uint8_t)(127.5 + 127.4 * sin(0.07 * y + 1.0));
    a.b = (uint8_t)(127.5 + 127.4 * sin(0.03 * (x + y) + 2.0));
    a.a = 255;
    return a;
}

__global__ void __launch_bounds__(256) synth_cloud_kernel(float4 *pos, pt_attr *attrs, size_t n,
                                                          pt_synth_params sp)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t ctr = sp.first_index + i;
    uint32_t r0[4], r1[4];
    philox4x32_10(ctr, 0, sp.seed, r0);
    philox4x32_10(ctr, 1, sp.seed, r1);
    double x, y, z, nx, ny, nz;
    if (sp.kind == PT_SYNTH_HEIGHTFIELD) {
        double u = sp.u0 + (sp.u1 - sp.u0) * u01(r0[0], r0[1]);
        double v = sp.v0 + (sp.v1 - sp.v0) * u01(r0[2], r0[3]);
        double zs, zx, zy;
        surface(u, v, zs, zx, zy);
        double inv = 1.0 / sqrt(zx * zx + zy * zy + 1.0);
        nx = -zx * inv; ny = -zy * inv; nz = inv;
        double g0, g1;
        gauss2(r1[0], r1[1], r1[2], r1[3], g0, g1);
        double off = sp.sigma * g0;
        x = u + off * nx; y = v + off * ny; z = zs + off * nz;
    } else {
        // skewed: 90% of the points in 64 Gaussian clusters whose centres lie on the surface over
        // the first 80% of the u-range, 10% uniform over that same part; the last 20% stays empty.
        double uspan = 0.8 * (sp.u1 - sp.u0);
        bool clustered = (r1[3] % 10u) != 0u;
        double cu, cv;
        if (clustered) {
            uint32_t cid = r1[2] & 63u;
            uint32_t rc[4];
            philox4x32_10(cid, 7, sp.seed, rc);
            cu = sp.u0 + uspan * u01(rc[0], rc[1]);
            cv = sp.v0 + (sp.v1 - sp.v0) * u01(rc[2], rc[3]);
        } else {
            cu = sp.u0 + uspan * u01(r0[0], r0[1]);
            cv = sp.v0 + (sp.v1 - sp.v0) * u01(r0[2], r0[3]);
        }
        double zs, zx, zy;
        surface(cu, cv, zs, zx, zy);
        double inv = 1.0 / sqrt(zx * zx + zy * zy + 1.0);
        nx = -zx * inv; ny = -zy * inv; nz = inv;
        x = cu; y = cv; z = zs;
        if (clustered) {
            uint32_t r2[4];
            philox4x32_10(ctr, 2, sp.seed, r2);
            double g0, g1, g2, g3;
            gauss2(r0[0], r0[1], r0[2], r0[3], g0, g1);
            gauss2(r2[0], r2[1], r2[2], r2[3], g2, g3);
            x += sp.sigma * g0; y += sp.sigma * g1; z += sp.sigma * g2;
        }
    }
    pos[i] = make_float4((float)x, (float)y, (float)z, 0.0f);
    if (attrs) attrs[i] = make_attr(x, y, nx, ny, nz);
}

__global__ void __launch_bounds__(256) synth_samples_kernel(double *q, size_t gu, size_t gv,
                                                            double u0, double u1, double v0,
                                                            double v1, int center)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= gu * gv) return;
    size_t i = t % gu, j = t / gu;
    double fu = center ? ((double)i + 0.5) / (double)gu : (gu > 1 ? (double)i / (double)(gu - 1) : 0.5);
    double fv = center ? ((double)j + 0.5) / (double)gv : (gv > 1 ? (double)j / (double)(gv - 1) : 0.5);
    double u = u0 + (u1 - u0) * fu, v = v0 + (v1 - v0) * fv;
    double z, zx, zy;
    surface(u, v, z, zx, zy);
    q[3 * t] = (double)(float)u;
    q[3 * t + 1] = (double)(float)v;
    q[3 * t + 2] = (double)(float)z;
}

struct Raw80s {
    double ver[3];
    double normal[3];
    int    color[3];
    int    pad;
    double U, V;
};

__global__ void __launch_bounds__(256) pack_points_kernel(const float4 *pos, const pt_attr *attrs,
                                                          size_t n, Raw80s *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos[i];
    Raw80s r;
    r.ver[0] = p.x; r.ver[1] = p.y; r.ver[2] = p.z;
    if (attrs) {
        pt_attr a = attrs[i];
        r.normal[0] = a.nx; r.normal[1] = a.ny; r.normal[2] = a.nz;
        r.color[0] = a.r; r.color[1] = a.g; r.color[2] = a.b;
    } else {
        r.normal[0] = r.normal[1] = r.normal[2] = 0.0;
        r.color[0] = r.color[1] = r.color[2] = 0;
    }
    r.pad = 0; r.U = 0.0; r.V = 0.0;
    out[i] = r;
}

__global__ void __launch_bounds__(256) pack_queries_kernel(const double *q, size_t m, Raw80s *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    Raw80s r;
    r.ver[0] = q[3 * i]; r.ver[1] = q[3 * i + 1]; r.ver[2] = q[3 * i + 2];
    r.normal[0] = r.normal[1] = r.normal[2] = 0.0;
    r.color[0] = r.color[1] = r.color[2] = 0;
    r.pad = 0; r.U = 0.0; r.V = 0.0;
    out[i] = r;
}

}  // namespace pt

using namespace pt;

static inline unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }

extern "C" {

int pt_synth_cloud_device(float *pos, pt_attr *attrs, size_t n, const pt_synth_params *params,
                          void *stream)
{
    if ((!pos && n) || !params) return PT_ERR_INVALID_ARG;
    if (n == 0) return PT_OK;
    synth_cloud_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>((float4 *)pos, attrs, n,
                                                                        *params);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_synth_samples_device(double *queries_xyz, size_t gu, size_t gv, double u0, double u1,
                            double v0, double v1, int center, void *stream)
{
    if (!queries_xyz && gu != 0 && gv != 0) return PT_ERR_INVALID_ARG;
    if (gu * gv == 0) return PT_OK;
    synth_samples_kernel<<<blocks_for(gu * gv), 256, 0, (cudaStream_t)stream>>>(
        queries_xyz, gu, gv, u0, u1, v0, v1, center);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_synth_pack_points_device(const float *pos, const pt_attr *attrs, size_t n, void *points80,
                                void *stream)
{
    if ((!pos || !points80) && n) return PT_ERR_INVALID_ARG;
    if (n == 0) return PT_OK;
    pack_points_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)pos, attrs, n, (Raw80s *)points80);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_synth_pack_queries_device(const double *queries_xyz, size_t m, void *points80, void *stream)
{
    if ((!queries_xyz || !points80) && m) return PT_ERR_INVALID_ARG;
    if (m == 0) return PT_OK;
    pack_queries_kernel<<<blocks_for(m), 256, 0, (cudaStream_t)stream>>>(queries_xyz, m,
                                                                         (Raw80s *)points80);
    count_launch();
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

}  // extern "C"
