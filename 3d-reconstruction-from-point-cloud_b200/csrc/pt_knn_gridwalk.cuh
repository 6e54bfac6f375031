// pt_knn_gridwalk.cuh -- variant 7 ("grid walker"): the scan kernel's per-thread top-k
// (pt_knn_scan.cuh) fed by the uniform-grid cell tables instead of the box-pyramid traversal.
//
// One THREAD per sample.  The warp-per-sample grid kernel (pt_knn_grid.cuh) spends ~85 % of its
// issue slots on warp-uniform bookkeeping that 32 lanes execute redundantly; here that
// bookkeeping is per-thread work of 32 different samples, at the price of uncoalesced run reads
// (every lane streams its own cell runs, 16-byte loads, fully consumed).  Same index, same
// termination test (Distance::min_distance_to_rectangle, /root/reference src/Distance.h:27-57, on
// the searched block), same exactness rules as the scan kernel:
//   * cells of the (2 rc + 1)^3 block are visited nearest-first; once the list is full a cell
//     whose lower bound (lattice units, rounded toward the safe side) exceeds the k-th distance
//     is skipped without a look-up;
//   * a block that is exhausted without proof starts the next attempt of the launch's schedule
//     from an EMPTY candidate list (a larger block contains the smaller one);
//   * dense buckets, runs above GW_MAX_RUN points and an exhausted schedule hand the sample over
//     to the box-pyramid kernels.
#pragma once

namespace pt {

constexpr uint32_t GW_MAX_RUN = 4096;

// the 125 cells of the 5^3 block sorted by squared distance from the centre (the first 27 are the
// 3^3 block): (dx + 2) | (dy + 2) << 3 | (dz + 2) << 6
__device__ const uint16_t g_cell_order[125] = {
    146, 145, 147, 138, 154, 82, 210, 137, 153, 139, 155, 81, 209, 83, 211, 74,
    202, 90, 218, 73, 201, 89, 217, 75, 203, 91, 219, 144, 148, 130, 162, 18,
    274, 136, 152, 140, 156, 129, 161, 131, 163, 80, 208, 84, 212, 66, 194, 98,
    226, 17, 273, 19, 275, 10, 266, 26, 282, 72, 200, 88, 216, 76, 204, 92,
    220, 65, 193, 97, 225, 67, 195, 99, 227, 9, 265, 25, 281, 11, 267, 27,
    283, 128, 160, 132, 164, 16, 272, 20, 276, 2, 258, 34, 290, 64, 192, 96,
    224, 68, 196, 100, 228, 8, 264, 24, 280, 12, 268, 28, 284, 1, 257, 33,
    289, 3, 259, 35, 291, 0, 256, 32, 288, 4, 260, 36, 292,
};

template <typename PT>
struct GridWalker {
    const QueryParams &P;
    double t[3];                 // lattice position of the sample (not clamped)
    uint32_t c21[3];             // lattice cell (clamped), the arithmetic of morton_kernel
    double slack;
    int attempt = 0, n = 0, n_cells = 0, rc = 0, sh = 0;
    uint32_t cc[3] = {0, 0, 0}, ncell = 0, cap = 0;
    const GridBucket *buckets = nullptr;
    float g2f = 0.f;             // squared guaranteed radius of the block, real units, rounded down
    float fdn[3], fup[3];        // position inside the own cell, lattice units, bracketed
    float S = 0.f, inv2f = 0.f;  // cell size in lattice units; (lattice units per real unit)^2, rounded up
    bool failed = false, restart = false;

    __device__ __forceinline__ GridWalker(const QueryParams &p, double qx, double qy, double qz, bool valid)
        : P(p)
    {
        const GridParams &G = p.grid;
        t[0] = (qx - G.lo[0]) * G.inv_cell21;
        t[1] = (qy - G.lo[1]) * G.inv_cell21;
        t[2] = (qz - G.lo[2]) * G.inv_cell21;
#pragma unroll
        for (int a = 0; a < 3; ++a) c21[a] = min(__double2uint_rz(t[a]), 2097151u);
        slack = G.slack + 1.8e-15 * fmax(fmax(fabs(qx), fabs(qy)), fabs(qz));
        inv2f = __double2float_ru(__dmul_ru(G.inv_cell21, G.inv_cell21));
        if (valid && G.n_attempts > 0) begin(0);
    }

    __device__ __forceinline__ void begin(int a)
    {
        const GridParams &G = P.grid;
        attempt = a;
        const GridTable T = G.tab[G.att_tab[a]];
        rc = G.att_rc[a];
        sh = 21 - T.level;
        ncell = 1u << T.level;
        buckets = T.buckets;
        cap = T.cap;
        n = 0;
        n_cells = rc == 1 ? 27 : 125;
        S = (float)(1u << sh);
        const float below = (float)rc * S, above = (float)(rc + 1) * S;
        float gl = INFINITY;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            cc[ax] = c21[ax] >> sh;
            const double f = t[ax] - (double)(cc[ax] << sh);
            fdn[ax] = __double2float_rd(f);
            fup[ax] = __double2float_ru(f);
            const float dl = __fadd_rd(fdn[ax], below);
            const float du = __fsub_rd(above, fup[ax]);
            gl = fminf(gl, cc[ax] > (uint32_t)rc ? dl : INFINITY);
            gl = fminf(gl, cc[ax] + rc + 1 < ncell ? du : INFINITY);
        }
        const double g = (double)gl * G.cell21 - slack;
        g2f = g > 0.0 ? __double2float_rd(__dmul_rd(g, g)) : 0.0f;
    }

    // lower bound (lattice units squared, rounded down) of the distance to cell offset d on one axis
    __device__ __forceinline__ float axis_gap(int ax, int d) const
    {
        if (d == 0) return 0.0f;
        const float e = d > 0 ? __fsub_rd((float)d * S, fup[ax]) : __fadd_rd(fdn[ax], (float)(-d - 1) * S);
        return fmaxf(e, 0.0f);
    }

    // The next non-empty cell run [rb, re) worth scanning under `bound`; false when the lane has
    // none this round (done, or a new attempt begins: take_restart()).
    __device__ __forceinline__ bool next_run(float bound, bool filling, bool &done, uint32_t &rb, uint32_t &re)
    {
        while (!done) {
            if (n == n_cells) {
                if (bound <= g2f) { done = true; break; }                 // the block proves the list
                if (attempt + 1 < P.grid.n_attempts) { begin(attempt + 1); restart = true; break; }
                failed = true; done = true; break;
            }
            const unsigned code = g_cell_order[n++];
            const int dx = (int)(code & 7u) - 2, dy = (int)((code >> 3) & 7u) - 2, dz = (int)(code >> 6) - 2;
            const uint32_t x = cc[0] + (uint32_t)dx, y = cc[1] + (uint32_t)dy, z = cc[2] + (uint32_t)dz;
            if (x >= ncell || y >= ncell || z >= ncell) continue;       // unsigned: also "negative" cells
            if (!filling) {
                const float ex = axis_gap(0, dx), ey = axis_gap(1, dy), ez = axis_gap(2, dz);
                const float lb = __fadd_rd(__fadd_rd(__fmul_rd(ex, ex), __fmul_rd(ey, ey)), __fmul_rd(ez, ez));
                if (lb > __fmul_ru(bound, inv2f)) continue;              // nothing in that cell can enter the list
            }
            uint32_t start, cnt;
            grid_lookup(buckets, cap, x, y, z, start, cnt);
            if (cnt == 0) continue;
            if (cnt > GW_MAX_RUN) { failed = true; done = true; break; }  // dense: the box pyramid's job
            rb = start;
            re = start + cnt;
            return true;
        }
        return false;
    }

    __device__ __forceinline__ bool take_restart()
    {
        const bool r = restart;
        restart = false;
        return r;
    }
};

}  // namespace pt
