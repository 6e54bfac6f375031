/*
 * pt_synth.h -- synthetic workload generators (bench / test harness only).
 *
 * The reference ships no sample clouds (*.ply is git-ignored,
 * /root/reference/.gitignore:5), so BASELINE.json's configs are synthetic
 * (SURVEY.md section 8 row M1): a noisy heightfield scan or a skewed-density
 * cluster cloud, generated on the device from a counter-based RNG
 * (Philox4x32-10, key = seed, counter = global point index) so every rank can
 * regenerate its own slab.  All coordinates are fp32-representable.
 *
 * Not part of the reference-facing ABI (include/points_transfer.h).
 */
#ifndef PT_SYNTH_H
#define PT_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#include "points_transfer.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PT_SYNTH_HEIGHTFIELD 0 /* z = sum a_i sin(f_i x+p_i) sin(g_i y+q_i) + N(0, sigma) along the normal */
#define PT_SYNTH_SKEWED      1 /* 90% in 64 Gaussian clusters, 10% uniform, last 20% of the u-range empty */

typedef struct pt_synth_params {
    int      kind;        /* PT_SYNTH_* */
    uint64_t seed;
    uint64_t first_index; /* global index of this slab's first point (RNG counter base) */
    double   u0, u1;      /* slab range along x */
    double   v0, v1;      /* range along y */
    double   sigma;       /* scan noise (heightfield) / cluster sigma (skewed) */
} pt_synth_params;

/* pos: n float4 (x,y,z,0) device; attrs: n pt_attr device (may be NULL). */
int pt_synth_cloud_device(float *pos, pt_attr *attrs, size_t n,
                          const pt_synth_params *params, void *stream);

/* gu x gv samples on the noise-free surface, row-major (v outer, u inner):
 * queries_xyz[(j*gu+i)*3 + {0,1,2}] doubles holding fp32-representable values.
 * center != 0 places samples at texel centres ((i+.5)/gu), else on the closed
 * vertex grid (i/(gu-1)). */
int pt_synth_samples_device(double *queries_xyz, size_t gu, size_t gv, double u0,
                            double u1, double v0, double v1, int center,
                            void *stream);

/* Expand device SoA (float4 pos + attrs) into host-layout 80-byte Point records
 * on the device (so they can be copied out for the host API / the oracle). */
int pt_synth_pack_points_device(const float *pos, const pt_attr *attrs, size_t n,
                                void *points80, void *stream);
/* Same for m*3 double query coordinates (normal/colour/UV zeroed). */
int pt_synth_pack_queries_device(const double *queries_xyz, size_t m,
                                 void *points80, void *stream);

#ifdef __cplusplus
}
#endif
#endif
