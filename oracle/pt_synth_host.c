/*
 * pt_synth_host.c -- host restatement of the synthetic workload generators
 * (csrc/pt_synth.cu of the product's bench-only library).
 *
 * TEST / BENCH INFRASTRUCTURE ONLY, like the rest of oracle/ (see pt_oracle.h): it lets
 * bench.py's `--impl reference` arm and the cpu_baseline leg generate BASELINE.json's clouds at
 * full size on the host cores without loading any CUDA code.  Same Philox4x32-10 streams (key =
 * seed, counter = global point index) and the same formulas in the same operation order as the
 * device generator; the only difference is the host libm (sin / cos / log differ from the
 * device's in the last ulp), so after the final rounding to fp32 a few coordinates per million
 * may differ by one fp32 ulp (tests/test_gpu_parity.py::test_host_generator_matches_device).
 */
#include "pt_oracle.h"

#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void philox4x32_10(uint64_t counter, uint32_t stream, uint64_t seed, uint32_t out[4])
{
    uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = stream, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u01(uint32_t a, uint32_t b)
{
    return (double)((((uint64_t)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}

static void surface(double x, double y, double *z, double *zx, double *zy)
{
    static const double A[4] = {12.0, 6.0, 2.5, 0.8}, F[4] = {0.021, 0.047, 0.11, 0.31},
                        G[4] = {0.017, 0.039, 0.13, 0.27}, Pp[4] = {0.3, 1.7, 2.9, 0.5},
                        Qp[4] = {1.1, 0.2, 4.1, 3.3};
    *z = 0; *zx = 0; *zy = 0;
    for (int i = 0; i < 4; ++i) {
        const double sx = sin(F[i] * x + Pp[i]), cx = cos(F[i] * x + Pp[i]);
        const double sy = sin(G[i] * y + Qp[i]), cy = cos(G[i] * y + Qp[i]);
        *z += A[i] * sx * sy;
        *zx += A[i] * F[i] * cx * sy;
        *zy += A[i] * G[i] * sx * cy;
    }
}

static void gauss2(const uint32_t r[4], double *g0, double *g1)
{
    const double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
    const double rad = sqrt(-2.0 * log(1.0 - u1));
    *g0 = rad * cos(6.283185307179586 * u2);
    *g1 = rad * sin(6.283185307179586 * u2);
}

int pto_synth_cloud(pto_point *out, int64_t n, int kind, uint64_t seed, uint64_t first_index,
                    double u0, double u1, double v0, double v1, double sigma, int nthreads)
{
    if (!out && n) return 1;
    if (nthreads <= 0) nthreads = pto_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t ctr = first_index + (uint64_t)i;
        uint32_t r0[4], r1[4];
        philox4x32_10(ctr, 0, seed, r0);
        philox4x32_10(ctr, 1, seed, r1);
        double x, y, z, nx, ny, nz;
        if (kind == 0) {   /* PT_SYNTH_HEIGHTFIELD */
            const double u = u0 + (u1 - u0) * u01(r0[0], r0[1]);
            const double v = v0 + (v1 - v0) * u01(r0[2], r0[3]);
            double zs, zx, zy, g0, g1;
            surface(u, v, &zs, &zx, &zy);
            const double inv = 1.0 / sqrt(zx * zx + zy * zy + 1.0);
            nx = -zx * inv; ny = -zy * inv; nz = inv;
            gauss2(r1, &g0, &g1);
            const double off = sigma * g0;
            x = u + off * nx; y = v + off * ny; z = zs + off * nz;
        } else {           /* PT_SYNTH_SKEWED */
            const double uspan = 0.8 * (u1 - u0);
            const int clustered = (r1[3] % 10u) != 0u;
            double cu, cv, zs, zx, zy;
            if (clustered) {
                uint32_t rc[4];
                philox4x32_10(r1[2] & 63u, 7, seed, rc);
                cu = u0 + uspan * u01(rc[0], rc[1]);
                cv = v0 + (v1 - v0) * u01(rc[2], rc[3]);
            } else {
                cu = u0 + uspan * u01(r0[0], r0[1]);
                cv = v0 + (v1 - v0) * u01(r0[2], r0[3]);
            }
            surface(cu, cv, &zs, &zx, &zy);
            const double inv = 1.0 / sqrt(zx * zx + zy * zy + 1.0);
            nx = -zx * inv; ny = -zy * inv; nz = inv;
            x = cu; y = cv; z = zs;
            if (clustered) {
                uint32_t r2[4];
                double g0, g1, g2, g3;
                philox4x32_10(ctr, 2, seed, r2);
                gauss2(r0, &g0, &g1);
                gauss2(r2, &g2, &g3);
                x += sigma * g0; y += sigma * g1; z += sigma * g2;
            }
        }
        pto_point *p = &out[i];
        p->ver[0] = (double)(float)x; p->ver[1] = (double)(float)y; p->ver[2] = (double)(float)z;
        p->normal[0] = (double)(float)nx; p->normal[1] = (double)(float)ny; p->normal[2] = (double)(float)nz;
        p->color[0] = (int)(uint8_t)(127.5 + 127.4 * sin(0.05 * x));
        p->color[1] = (int)(uint8_t)(127.5 + 127.4 * sin(0.07 * y + 1.0));
        p->color[2] = (int)(uint8_t)(127.5 + 127.4 * sin(0.03 * (x + y) + 2.0));
        p->pad_ = 0; p->U = 0.0; p->V = 0.0;
    }
    return 0;
}

int pto_synth_samples(pto_point *out, int64_t gu, int64_t gv, double u0, double u1, double v0,
                      double v1, int center)
{
    if (!out && gu > 0 && gv > 0) return 1;
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < gu * gv; ++t) {
        const int64_t i = t % gu, j = t / gu;
        const double fu = center ? ((double)i + 0.5) / (double)gu : (gu > 1 ? (double)i / (double)(gu - 1) : 0.5);
        const double fv = center ? ((double)j + 0.5) / (double)gv : (gv > 1 ? (double)j / (double)(gv - 1) : 0.5);
        const double u = u0 + (u1 - u0) * fu, v = v0 + (v1 - v0) * fv;
        double z, zx, zy;
        surface(u, v, &z, &zx, &zy);
        pto_point *p = &out[t];
        p->ver[0] = (double)(float)u; p->ver[1] = (double)(float)v; p->ver[2] = (double)(float)z;
        p->normal[0] = p->normal[1] = p->normal[2] = 0.0;
        p->color[0] = p->color[1] = p->color[2] = 0;
        p->pad_ = 0; p->U = 0.0; p->V = 0.0;
    }
    return 0;
}
