// pt_texture.cu -- per-face transfer and texture output on the GPU (SURVEY.md section 8 rows
// N1 / N3 / N4): everything the reference does AFTER the neighbour search,
// /root/reference src/pointsTransfer.cpp:
//   :484-537  project the neighbours of a face's corners onto the face plane, keep those inside
//   :539-581  2-D Delaunay triangulation of corners + inside points, UVs by barycentric mix
//   :66-107   draw_triangle -- rasterise every sub-triangle into the BGRA texture
//   :593-611  25 x 25 dilate, gutter = dilated & ~alpha, padded = texture + gutter
//
// One thread per face runs :484-581 start to finish (the data of a face -- <= 3 K neighbours --
// lives in its local arrays) and rasterises its sub-triangles itself.  The reference draws faces
// one after the other, later ones overwriting earlier ones where sub-triangles meet; here every
// pixel write is a 64-bit atomicMax on (face << 8 | sub-triangle rank) << 32 | BGRA, so the
// result is that same "last writer in reference order wins", independent of scheduling.
//
// CGAL is restated from memory (flagged in oracle/pt_texture_oracle.c, the CPU statement of the
// same definitions, which the tests compare this file with bit for bit): Plane_3::to_2d's
// stretched frame, Triangle_coordinates_2, and Delaunay by its definition (a non-degenerate
// triple is a face iff no other point is strictly inside its circumcircle; the in-circle sign is
// evaluated once per index-sorted quadruple).  Not reproduced: the out-of-bounds row
// `resolution - j` at j = 0 / column `resolution` (:100-103) and the uninitialised Mat (:402).
#include <cmath>

#include "pt_index.cuh"

namespace pt {

constexpr int TEX_MAX_NB = 3 * PT_MAX_K;      // neighbours of one face before the union
constexpr int TEX_MAX_PTS = 3 + TEX_MAX_NB;

struct V2 { double a, b; };

struct MeshVertex {      // what a face needs of a mesh vertex (unpacked from the 80-byte Point)
    double x, y, z, u, v;
    int    r, g, b;
    int    pad;
};

struct Raw80t {
    double ver[3];
    double normal[3];
    int    color[3];
    int    pad;
    double U, V;
};

__global__ void __launch_bounds__(256)
tex_unpack_vertices_kernel(const Raw80t *raw, uint32_t n, MeshVertex *mv, double *xyz)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Raw80t r = raw[i];
    MeshVertex m;
    m.x = r.ver[0]; m.y = r.ver[1]; m.z = r.ver[2];
    m.u = r.U; m.v = r.V;
    m.r = r.color[0]; m.g = r.color[1]; m.b = r.color[2];
    m.pad = 0;
    mv[i] = m;
    xyz[3 * (size_t)i] = r.ver[0]; xyz[3 * (size_t)i + 1] = r.ver[1]; xyz[3 * (size_t)i + 2] = r.ver[2];
}

template <typename PT>
__global__ void __launch_bounds__(256) tex_inverse_perm_kernel(const PT *pts, uint32_t n, uint32_t *inv)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[pts[i].idx] = i;
}

__device__ __forceinline__ double tx_area2(V2 p, V2 q, V2 r)
{
    return (q.a - p.a) * (r.b - p.b) - (q.b - p.b) * (r.a - p.a);
}
__device__ __forceinline__ void tx_tri_coords(V2 v0, V2 v1, V2 v2, V2 x, double bc[3])
{
    const double inv = 1.0 / tx_area2(v0, v1, v2);
    bc[0] = tx_area2(v1, v2, x) * inv;
    bc[1] = tx_area2(v2, v0, x) * inv;
    bc[2] = 1.0 - bc[0] - bc[1];
}
__device__ __forceinline__ double tx_incircle(V2 a, V2 b, V2 c, V2 d)
{
    const double ax = a.a - d.a, ay = a.b - d.b, bx = b.a - d.a, by = b.b - d.b, cx = c.a - d.a, cy = c.b - d.b;
    const double al = ax * ax + ay * ay, bl = bx * bx + by * by, cl = cx * cx + cy * cy;
    return ax * (by * cl - bl * cy) - ay * (bx * cl - bl * cx) + al * (bx * cy - by * cx);
}
__device__ __forceinline__ double tx_dot3(const double *u, const double *v)
{
    return u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
}
__device__ __forceinline__ V2 tx_to_2d(const double x[3], const double r[3], const double n[3],
                                       const double b1[3], const double b2[3])
{
    const double w[3] = {x[0] - r[0], x[1] - r[1], x[2] - r[2]};
    const double t = tx_dot3(w, n) / tx_dot3(n, n);
    const double pr[3] = {w[0] - t * n[0], w[1] - t * n[1], w[2] - t * n[2]};
    V2 o;
    o.a = tx_dot3(pr, b1) / tx_dot3(b1, b1);
    o.b = tx_dot3(pr, b2) / tx_dot3(b2, b2);
    return o;
}

struct TexVertex { double u, v; int r, g, b; };

// src/pointsTransfer.cpp:66-107; `order` = (face << 8 | sub-triangle rank), later wins
__device__ void tx_draw_triangle(const TexVertex &t0, const TexVertex &t1, const TexVertex &t2, int res,
                                 unsigned long long *canvas, unsigned long long order)
{
    V2 p, q, r;
    p.a = t0.u * res; p.b = t0.v * res;
    q.a = t1.u * res; q.b = t1.v * res;
    r.a = t2.u * res; r.b = t2.v * res;
    const double xmin = fmin(p.a, fmin(q.a, r.a)), xmax = fmax(p.a, fmax(q.a, r.a));
    const double ymin = fmin(p.b, fmin(q.b, r.b)), ymax = fmax(p.b, fmax(q.b, r.b));
    if (!(xmax - xmin < 4.0 * res) || !(ymax - ymin < 4.0 * res)) return;      // NaN / absurd UVs
    const double fx1 = floor(xmax), fy1 = floor(ymax);
    for (int i = (int)floor(xmin); i <= fx1; i++) {
        for (int j = (int)floor(ymin); j <= fy1; j++) {
            int x = i, y = j;
            if (x >= res) x = res - 1;
            if (y >= res) y = res - 1;
            double bc[3];
            V2 px;
            px.a = (double)x; px.b = (double)y;
            tx_tri_coords(p, q, r, px, bc);
            if (bc[0] >= 0 && bc[1] >= 0 && bc[2] >= 0) {
                const float fr = (float)(bc[0] * t0.r + bc[1] * t1.r + bc[2] * t2.r);
                const float fg = (float)(bc[0] * t0.g + bc[1] * t1.g + bc[2] * t2.g);
                const float fb = (float)(bc[0] * t0.b + bc[1] * t1.b + bc[2] * t2.b);
                const int row = res - j, col = i;
                if (row < 0 || row >= res || col < 0 || col >= res) continue;    // out of bounds in the reference
                const unsigned bgra = ((unsigned)(int)fb & 0xffu) | (((unsigned)(int)fg & 0xffu) << 8) |
                                      (((unsigned)(int)fr & 0xffu) << 16) | 0xff000000u;
                atomicMax(canvas + (size_t)row * res + col, (order << 32) | bgra);
            }
        }
    }
}

// where the position of cloud point `id` (original index) comes from
template <typename PT> struct PosSorted {      // an index: sorted records + inverse permutation
    const PT *pts;
    const uint32_t *inv;
    __device__ __forceinline__ void get(int id, double x[3]) const
    {
        const PT rec = pts[inv[id]];
        x[0] = (double)rec.x; x[1] = (double)rec.y; x[2] = (double)rec.z;
    }
};
struct PosPlain {                               // n x 3 doubles in original order
    const double *xyz;
    __device__ __forceinline__ void get(int id, double x[3]) const
    {
        x[0] = xyz[3 * (size_t)id]; x[1] = xyz[3 * (size_t)id + 1]; x[2] = xyz[3 * (size_t)id + 2];
    }
};

template <typename POS>
__global__ void __launch_bounds__(64)
tex_face_kernel(const POS pos, uint32_t n_points, const pt_attr *attrs, const MeshVertex *mv,
                const int32_t *idx, int k, const int32_t *faces, uint32_t n_faces, uint32_t n_vertices,
                int res, unsigned long long *canvas, unsigned long long *stats)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    int vi[3];
    MeshVertex tv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        vi[c] = faces[3 * (size_t)f + c];
        if (vi[c] < 0 || (uint32_t)vi[c] >= n_vertices) return;      // malformed face: nothing to draw
        tv[c] = mv[vi[c]];
    }
    // union of the corners' neighbour lists (src/pointsTransfer.cpp:470-479), by point index
    int nb[TEX_MAX_NB], n_nb = 0;
    for (int c = 0; c < 3; ++c)
        for (int j = 0; j < k; ++j) {
            const int id = idx[(size_t)vi[c] * k + j];
            if (id < 0 || (uint32_t)id >= n_points) continue;
            bool seen = false;
            for (int e = 0; e < n_nb; ++e) seen |= nb[e] == id;
            if (!seen) nb[n_nb++] = id;
        }
    // Plane_3(r, p, q) and its 2-D frame (:490-494)
    const double R[3] = {tv[0].x, tv[0].y, tv[0].z};
    const double e1[3] = {tv[1].x - R[0], tv[1].y - R[1], tv[1].z - R[2]};
    const double e2[3] = {tv[2].x - R[0], tv[2].y - R[1], tv[2].z - R[2]};
    double n[3], b1[3], b2[3];
    n[0] = e1[1] * e2[2] - e1[2] * e2[1];
    n[1] = e1[2] * e2[0] - e1[0] * e2[2];
    n[2] = e1[0] * e2[1] - e1[1] * e2[0];
    {
        const double a = n[0], b = n[1], c = n[2];
        if (a == 0.0) { b1[0] = 1; b1[1] = 0; b1[2] = 0; }
        else if (b == 0.0) { b1[0] = 0; b1[1] = 1; b1[2] = 0; }
        else if (c == 0.0) { b1[0] = 0; b1[1] = 0; b1[2] = 1; }
        else if (fabs(a) <= fabs(b) && fabs(a) <= fabs(c)) { b1[0] = 0; b1[1] = -c; b1[2] = b; }
        else if (fabs(b) <= fabs(a) && fabs(b) <= fabs(c)) { b1[0] = -c; b1[1] = 0; b1[2] = a; }
        else { b1[0] = -b; b1[1] = a; b1[2] = 0; }
    }
    b2[0] = n[1] * b1[2] - n[2] * b1[1];
    b2[1] = n[2] * b1[0] - n[0] * b1[2];
    b2[2] = n[0] * b1[1] - n[1] * b1[0];

    V2 P[TEX_MAX_PTS];
    TexVertex V[TEX_MAX_PTS];
    int np = 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double x[3] = {tv[c].x, tv[c].y, tv[c].z};
        P[c] = tx_to_2d(x, R, n, b1, b2);
        V[c].u = tv[c].u; V[c].v = tv[c].v;
        V[c].r = tv[c].r; V[c].g = tv[c].g; V[c].b = tv[c].b;
    }
    for (int e = 0; e < n_nb; ++e) {
        double x[3];
        pos.get(nb[e], x);
        const V2 xp = tx_to_2d(x, R, n, b1, b2);
        double bc[3];
        tx_tri_coords(P[0], P[1], P[2], xp, bc);
        if (bc[0] >= 0 && bc[1] >= 0 && bc[2] >= 0) {        // :530
            const pt_attr at = attrs[nb[e]];
            P[np] = xp;
            V[np].u = bc[0] * tv[0].u + bc[1] * tv[1].u + bc[2] * tv[2].u;      // :572
            V[np].v = bc[0] * tv[0].v + bc[1] * tv[1].v + bc[2] * tv[2].v;      // :573
            V[np].r = at.r; V[np].g = at.g; V[np].b = at.b;
            ++np;
        }
    }
    unsigned drawn = 0;
    const unsigned long long fkey = (unsigned long long)f << 8;
    if (np == 3) {                                           // :540-544
        tx_draw_triangle(V[0], V[1], V[2], res, canvas, fkey);
        drawn = 1;
    } else {                                                 // :545-581, Delaunay by its definition
        for (int a = 0; a < np; ++a)
            for (int b = a + 1; b < np; ++b)
                for (int c = b + 1; c < np; ++c) {
                    const double o = tx_area2(P[a], P[b], P[c]);
                    if (o == 0.0 || o != o) continue;
                    bool ok = true;
                    for (int d = 0; d < np && ok; ++d) {
                        if (d == a || d == b || d == c) continue;
                        // in-circle determinant of the index-sorted quadruple, signed by the
                        // permutation that sorts (a, b, c, d): see oracle/pt_texture_oracle.c
                        int q0 = a, q1 = b, q2 = c, q3 = d, sw = 0;
                        if (q3 < q2) { const int t = q3; q3 = q2; q2 = t; ++sw; }
                        if (q2 < q1) { const int t = q2; q2 = q1; q1 = t; ++sw; }
                        if (q1 < q0) { const int t = q1; q1 = q0; q0 = t; ++sw; }
                        double s = tx_incircle(P[q0], P[q1], P[q2], P[q3]);
                        if (sw & 1) s = -s;
                        if (o < 0) s = -s;
                        if (s > 0) ok = false;
                    }
                    if (!ok) continue;
                    tx_draw_triangle(V[a], V[b], V[c], res, canvas, fkey | (drawn < 255u ? drawn : 255u));
                    ++drawn;
                }
    }
    if (stats) {
        atomicAdd(&stats[0], (unsigned long long)drawn);
        atomicAdd(&stats[1], (unsigned long long)(np - 3));
    }
}

__global__ void __launch_bounds__(256)
tex_resolve_kernel(const unsigned long long *canvas, size_t n_pix, uint32_t *tex)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pix) tex[i] = (uint32_t)(canvas[i] & 0xffffffffull);     // never written: 0 (alpha 0)
}

// :593-611 as two separable 25-tap byte-wise maxima (OpenCV ignores the border) + the combine
__global__ void __launch_bounds__(256) tex_dilate_h_kernel(const uint32_t *tex, int res, uint32_t *tmp)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)res * res) return;
    const int y = (int)(i / res), x = (int)(i % res);
    uint32_t m = 0;
    for (int dx = -12; dx <= 12; ++dx) {
        const int xx = x + dx;
        if (xx >= 0 && xx < res) m = __vmaxu4(m, tex[(size_t)y * res + xx]);
    }
    tmp[i] = m;
}
__global__ void __launch_bounds__(256)
tex_dilate_v_pad_kernel(const uint32_t *tex, const uint32_t *tmp, int res, uint32_t *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)res * res) return;
    const int y = (int)(i / res), x = (int)(i % res);
    uint32_t m = 0;
    for (int dy = -12; dy <= 12; ++dy) {
        const int yy = y + dy;
        if (yy >= 0 && yy < res) m = __vmaxu4(m, tmp[(size_t)yy * res + x]);
    }
    const uint32_t t = tex[i];
    const uint32_t alpha = t >> 24;
    const uint32_t not_alpha = (alpha ^ 0xffu) * 0x01010101u;      // ~alpha replicated to the 4 channels
    out[i] = __vaddus4(t, m & not_alpha);
}

static inline unsigned tcdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// Mesh upload + (optionally) the neighbour search + face kernel + post-process.  `ix` gives the
// stream, the query kernels (when d_idx_given is null) and the attribute array.
template <typename POS>
static int texture_pipeline(pt_index *ix, POS pos, uint32_t n_points, const pt_attr *attrs, bool search,
                            const int32_t *idx_host, const void *vertices, size_t n_vertices,
                            const int32_t *faces, size_t n_faces, int k, double radius, int res, int pad,
                            uint8_t *bgra_out, pt_texture_stats *st)
{
    cudaStream_t s = ix->stream;
    const size_t n_pix = (size_t)res * res;
    Raw80t *d_raw = nullptr;
    MeshVertex *d_mv = nullptr;
    double *d_xyz = nullptr;
    int32_t *d_idx = nullptr, *d_faces = nullptr;
    unsigned long long *d_canvas = nullptr, *d_stats = nullptr;
    uint32_t *d_tex = nullptr, *d_tmp = nullptr, *d_out = nullptr;
    cudaEvent_t ev[4] = {};
    int rc = PT_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == PT_OK) rc = map_cuda_error(e); return e != cudaSuccess; };
    do {
        for (auto &e : ev) if (fail(cudaEventCreate(&e))) break;
        if (rc != PT_OK) break;
        if (pool_alloc((void **)&d_raw, sizeof(Raw80t) * n_vertices, s) || pool_alloc((void **)&d_mv, sizeof(MeshVertex) * n_vertices, s) ||
            pool_alloc((void **)&d_xyz, sizeof(double) * 3 * n_vertices, s) || pool_alloc((void **)&d_idx, sizeof(int32_t) * n_vertices * k, s) ||
            pool_alloc((void **)&d_faces, sizeof(int32_t) * 3 * n_faces, s) || pool_alloc((void **)&d_canvas, sizeof(unsigned long long) * n_pix, s) ||
            pool_alloc((void **)&d_stats, 16, s) || pool_alloc((void **)&d_tex, 4 * n_pix, s) ||
            (pad && (pool_alloc((void **)&d_tmp, 4 * n_pix, s) || pool_alloc((void **)&d_out, 4 * n_pix, s)))) {
            rc = PT_ERR_OUT_OF_MEMORY;
            break;
        }
        if (fail(cudaEventRecord(ev[0], s))) break;
        if (fail(cudaMemcpyAsync(d_raw, vertices, sizeof(Raw80t) * n_vertices, cudaMemcpyHostToDevice, s))) break;
        if (fail(cudaMemcpyAsync(d_faces, faces, sizeof(int32_t) * 3 * n_faces, cudaMemcpyHostToDevice, s))) break;
        if (fail(cudaMemsetAsync(d_canvas, 0, sizeof(unsigned long long) * n_pix, s))) break;
        if (fail(cudaMemsetAsync(d_stats, 0, 16, s))) break;
        if (n_vertices) {
            tex_unpack_vertices_kernel<<<tcdiv(n_vertices, 256), 256, 0, s>>>(d_raw, (uint32_t)n_vertices, d_mv, d_xyz);
            count_launch();
        }
        if (search) {
            // the neighbour search of every UNIQUE vertex (the reference searches 3 x per face, :474)
            QueryParams qp{};
            qp.pts = ix->pts; qp.attrs = ix->attrs; qp.ids = nullptr; qp.pyr = ix->pyr;
            qp.n = ix->n; qp.n_leaves = ix->n_leaves; qp.w_levels = ix->w_levels; qp.t_levels = ix->t_levels;
            qp.pq_cap = opt_queue_cap();
            qp.queries = d_xyz; qp.m = (uint32_t)n_vertices; qp.k = k;
            qp.r2 = (!(radius >= 0.0) || std::isinf(radius)) ? INFINITY : radius * radius;
            qp.idx_out = d_idx;
            if ((rc = launch_query(ix, qp, s)) != PT_OK) break;
        } else if (n_vertices) {
            if (fail(cudaMemcpyAsync(d_idx, idx_host, sizeof(int32_t) * n_vertices * k, cudaMemcpyHostToDevice, s))) break;
        }
        if (fail(cudaEventRecord(ev[1], s))) break;
        if (n_faces) {
            tex_face_kernel<POS><<<tcdiv(n_faces, 64), 64, 0, s>>>(pos, n_points, attrs, d_mv, d_idx, k, d_faces,
                                                                (uint32_t)n_faces, (uint32_t)n_vertices, res, d_canvas,
                                                                d_stats);
            count_launch();
        }
        tex_resolve_kernel<<<tcdiv(n_pix, 256), 256, 0, s>>>(d_canvas, n_pix, d_tex);
        count_launch();
        if (fail(cudaEventRecord(ev[2], s))) break;
        const uint32_t *result = d_tex;
        if (pad) {
            tex_dilate_h_kernel<<<tcdiv(n_pix, 256), 256, 0, s>>>(d_tex, res, d_tmp);
            tex_dilate_v_pad_kernel<<<tcdiv(n_pix, 256), 256, 0, s>>>(d_tex, d_tmp, res, d_out);
            count_launch(2);
            result = d_out;
        }
        if (fail(cudaEventRecord(ev[3], s))) break;
        if (fail(cudaGetLastError())) break;
        if (fail(cudaMemcpyAsync(bgra_out, result, 4 * n_pix, cudaMemcpyDeviceToHost, s))) break;
        unsigned long long h_stats[2] = {0, 0};
        if (fail(cudaMemcpyAsync(h_stats, d_stats, 16, cudaMemcpyDeviceToHost, s))) break;
        if (fail(cudaStreamSynchronize(s))) break;
        if (st) {
            st->triangles = h_stats[0];
            st->inside_points = h_stats[1];
            cudaEventElapsedTime(&st->knn_ms, ev[0], ev[1]);
            cudaEventElapsedTime(&st->draw_ms, ev[1], ev[2]);
            cudaEventElapsedTime(&st->pad_ms, ev[2], ev[3]);
        }
    } while (0);
    cudaStreamSynchronize(s);
    void *frees[] = {d_raw, d_mv, d_xyz, d_idx, d_faces, d_canvas, d_stats, d_tex, d_tmp, d_out};
    for (void *p : frees) pool_free(p, s);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    cudaGetLastError();
    return rc;
}

template <typename PT>
static int texture_render_index(pt_index *ix, const void *vertices, size_t n_vertices, const int32_t *faces,
                                size_t n_faces, int k, double radius, int res, int pad, uint8_t *bgra_out,
                                pt_texture_stats *st)
{
    if (!ix->inv_perm && ix->n) {            // original index -> position in the sorted cloud, built once
        PT_TRY(dev_alloc((void **)&ix->inv_perm, sizeof(uint32_t) * (size_t)ix->n));
        tex_inverse_perm_kernel<PT><<<tcdiv(ix->n, 256), 256, 0, ix->stream>>>((const PT *)ix->pts, ix->n, ix->inv_perm);
        count_launch();
    }
    PosSorted<PT> pos{(const PT *)ix->pts, ix->inv_perm};
    return texture_pipeline(ix, pos, ix->n, ix->attrs, true, nullptr, vertices, n_vertices, faces, n_faces, k, radius,
                            res, pad, bgra_out, st);
}

}  // namespace pt

using namespace pt;

extern "C" int pt_texture_render(pt_index *ix, const void *vertices, size_t n_vertices, const int32_t *faces,
                                 size_t n_faces, int k, double radius, int resolution, int pad,
                                 uint8_t *bgra_out, pt_texture_stats *stats)
{
    if (!ix || (!vertices && n_vertices) || (!faces && n_faces) || !bgra_out || resolution < 1 || resolution > 32768 ||
        n_vertices > 0x7ffffff0ull || n_faces >= (1ull << 24))
        return PT_ERR_INVALID_ARG;
    if (k < 1 || k > PT_MAX_K) return PT_ERR_UNSUPPORTED;
    if (ix->ids) return PT_ERR_UNSUPPORTED;           // slab indexes answer with global ids
    if (!ix->attrs && ix->n) return PT_ERR_INVALID_ARG;
    PT_CUDA(cudaSetDevice(ix->device));
    return ix->coord_f64 ? texture_render_index<PointD>(ix, vertices, n_vertices, faces, n_faces, k, radius, resolution, pad, bgra_out, stats)
                         : texture_render_index<PointF>(ix, vertices, n_vertices, faces, n_faces, k, radius, resolution, pad, bgra_out, stats);
}

extern "C" int pt_texture_render_lists(const void *points, size_t n, const void *vertices, size_t n_vertices,
                                       const int32_t *faces, size_t n_faces, const int32_t *idx, int k,
                                       int resolution, int pad, int device, uint8_t *bgra_out,
                                       pt_texture_stats *stats)
{
    if ((!points && n) || (!vertices && n_vertices) || (!faces && n_faces) || (!idx && n_vertices) || !bgra_out ||
        resolution < 1 || resolution > 32768 || n >= 0x7fffffffull || n_vertices > 0x7ffffff0ull || n_faces >= (1ull << 24))
        return PT_ERR_INVALID_ARG;
    if (k < 1 || k > PT_MAX_K) return PT_ERR_UNSUPPORTED;
    if (pt_device_count() == 0) return PT_ERR_NO_DEVICE;
    if (device < 0) PT_CUDA(cudaGetDevice(&device));
    PT_CUDA(cudaSetDevice(device));
    // a bare handle: a stream plus the cloud's positions and attributes in original order
    pt_index tmp;
    tmp.device = device;
    double *xyz = nullptr;
    bool representable = true;
    int rc = map_cuda_error(cudaStreamCreateWithFlags(&tmp.stream, cudaStreamNonBlocking));
    if (rc == PT_OK) rc = ingest_points_aos(&tmp, points, n, PT_COORD_AUTO, &xyz, &representable);
    if (rc == PT_OK) {
        PosPlain pos{xyz};
        rc = texture_pipeline(&tmp, pos, (uint32_t)n, tmp.attrs, false, idx, vertices, n_vertices, faces, n_faces, k,
                              -1.0, resolution, pad, bgra_out, stats);
    }
    if (tmp.stream) cudaStreamSynchronize(tmp.stream);
    dev_free(xyz);
    pool_free(tmp.attrs, tmp.stream);
    if (tmp.stream) cudaStreamSynchronize(tmp.stream);
    if (tmp.stream) cudaStreamDestroy(tmp.stream);
    cudaGetLastError();
    return rc;
}
