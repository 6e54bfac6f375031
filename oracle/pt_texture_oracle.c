/*
 * pt_texture_oracle.c -- CPU restatement of the reference's per-face transfer and texture output
 * (SURVEY.md section 8 rows N1, N3, N4).  TEST INFRASTRUCTURE ONLY, like the rest of oracle/.
 *
 * Follows /root/reference/src/pointsTransfer.cpp:
 *   :465-479  per face, the union of the K nearest cloud points of its 3 corners
 *   :484-537  Plane_3(r, p, q) through the corners, orthogonal projection of every neighbour,
 *             plane.to_2d, Triangle_coordinates_2, keep the points with all bc >= 0
 *   :539-581  no inside point -> draw the face; else 2-D Delaunay of corners + inside points,
 *             UV of an inside point = barycentric mix of the corner UVs (:569-574), draw every
 *             finite face
 *   :66-107   draw_triangle: bounding-box scan, barycentric colour mix in double, stored through
 *             float into uchar (truncation), alpha 255, pixel (row resolution - j, column i)
 *   :593-615  25x25 rectangular dilate, edges = dilated & ~alpha, padded = texture + edges
 *
 * What cannot be pinned on the reference (CGAL and OpenCV C++ are absent from this image, the
 * reference ships no test image): the geometry below restates CGAL from memory and is flagged
 * "(CGAL, recalled)".  tests/test_texture.py pins the post-process on OpenCV's own dilate /
 * bitwise / add through the cv2 wheel.
 *
 *   Plane_3::to_2d (CGAL, recalled): coordinates in the basis base1 = an axis-aligned vector
 *     orthogonal to the plane normal n = (p - r) x (q - r), base2 = n x base1.  |base2| =
 *     |n| |base1|, so the 2-D frame is NOT isometric -- the Delaunay triangulation is taken in
 *     that stretched frame, as in the reference.
 *   Triangle_coordinates_2 (CGAL, recalled): b0 = area(v1, v2, x) / area(v0, v1, v2),
 *     b1 = area(v2, v0, x) / area(v0, v1, v2), b2 = 1 - b0 - b1.
 *   Delaunay_triangulation_2: CGAL's incremental algorithm is replaced by the definition -- a
 *     triple is a face iff it is non-degenerate and no other point lies strictly inside its
 *     circumcircle.  The in-circle sign is evaluated once per index-sorted quadruple and reused
 *     with the permutation's parity, so the faces are consistent.  Identical to CGAL for points
 *     in general position; exactly co-circular quadruples yield both diagonals (overlapping
 *     faces with the same vertex data) where CGAL picks one.  Faces are drawn in lexicographic
 *     index order (CGAL's face order is an implementation detail; later faces overwrite).
 *   The reference's union is a std::set with a non-strict-weak comparator (src/Point.h:94-102):
 *     reverse insertion order, position duplicates mostly kept.  Neither affects the output:
 *     a duplicate is the same 2-D point with the same data, and the face set of a Delaunay
 *     triangulation does not depend on insertion order.  The union here is by point index.
 *   Quirks NOT reproduced: the write to row `resolution - j` for j = 0 and column i = resolution
 *     is out of bounds (:100-103) -- skipped; `texture` is never initialised (:402) -- zeroed.
 */
#include "pt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double a, b; } v2;

static void plane_frame(const double r[3], const double p[3], const double q[3], double n[3],
                        double b1[3], double b2[3])
{
    const double e1[3] = {p[0] - r[0], p[1] - r[1], p[2] - r[2]};
    const double e2[3] = {q[0] - r[0], q[1] - r[1], q[2] - r[2]};
    n[0] = e1[1] * e2[2] - e1[2] * e2[1];
    n[1] = e1[2] * e2[0] - e1[0] * e2[2];
    n[2] = e1[0] * e2[1] - e1[1] * e2[0];
    const double a = n[0], b = n[1], c = n[2];
    if (a == 0.0) { b1[0] = 1; b1[1] = 0; b1[2] = 0; }
    else if (b == 0.0) { b1[0] = 0; b1[1] = 1; b1[2] = 0; }
    else if (c == 0.0) { b1[0] = 0; b1[1] = 0; b1[2] = 1; }
    else if (fabs(a) <= fabs(b) && fabs(a) <= fabs(c)) { b1[0] = 0; b1[1] = -c; b1[2] = b; }
    else if (fabs(b) <= fabs(a) && fabs(b) <= fabs(c)) { b1[0] = -c; b1[1] = 0; b1[2] = a; }
    else { b1[0] = -b; b1[1] = a; b1[2] = 0; }
    b2[0] = n[1] * b1[2] - n[2] * b1[1];
    b2[1] = n[2] * b1[0] - n[0] * b1[2];
    b2[2] = n[0] * b1[1] - n[1] * b1[0];
}

static double dot3(const double *u, const double *v) { return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]; }

/* 2-D coordinates of the orthogonal projection of x onto the plane, frame origin at r */
static v2 to_2d(const double x[3], const double r[3], const double n[3], const double b1[3],
                const double b2[3])
{
    const double w[3] = {x[0] - r[0], x[1] - r[1], x[2] - r[2]};
    const double t = dot3(w, n) / dot3(n, n);
    const double pr[3] = {w[0] - t * n[0], w[1] - t * n[1], w[2] - t * n[2]};
    v2 o = {dot3(pr, b1) / dot3(b1, b1), dot3(pr, b2) / dot3(b2, b2)};
    return o;
}

static double area2(v2 p, v2 q, v2 r)      /* twice the signed area */
{
    return (q.a - p.a) * (r.b - p.b) - (q.b - p.b) * (r.a - p.a);
}

static void tri_coords(v2 v0, v2 v1, v2 v2_, v2 x, double bc[3])
{
    const double inv = 1.0 / area2(v0, v1, v2_);
    bc[0] = area2(v1, v2_, x) * inv;
    bc[1] = area2(v2_, v0, x) * inv;
    bc[2] = 1.0 - bc[0] - bc[1];
}

/* > 0: d strictly inside the circumcircle of the counter-clockwise triangle (a, b, c) */
static double incircle(v2 a, v2 b, v2 c, v2 d)
{
    const double ax = a.a - d.a, ay = a.b - d.b, bx = b.a - d.a, by = b.b - d.b, cx = c.a - d.a, cy = c.b - d.b;
    const double al = ax * ax + ay * ay, bl = bx * bx + by * by, cl = cx * cx + cy * cy;
    return ax * (by * cl - bl * cy) - ay * (bx * cl - bl * cx) + al * (bx * cy - by * cx);
}

typedef struct { double u, v; int c[3]; } tex_vertex;

/* src/pointsTransfer.cpp:66-107 */
static void draw_triangle(const tex_vertex t[3], int res, uint8_t *bgra)
{
    const v2 p = {t[0].u * res, t[0].v * res}, q = {t[1].u * res, t[1].v * res}, r = {t[2].u * res, t[2].v * res};
    const double xmin = fmin(p.a, fmin(q.a, r.a)), xmax = fmax(p.a, fmax(q.a, r.a));
    const double ymin = fmin(p.b, fmin(q.b, r.b)), ymax = fmax(p.b, fmax(q.b, r.b));
    if (!(xmax - xmin < 4.0 * res) || !(ymax - ymin < 4.0 * res)) return;      /* NaN / absurd UVs */
    for (int i = (int)floor(xmin); i <= floor(xmax); i++) {
        for (int j = (int)floor(ymin); j <= floor(ymax); j++) {
            int x = i, y = j;
            if (x >= res) x = res - 1;
            if (y >= res) y = res - 1;
            double bc[3];
            const v2 px = {(double)x, (double)y};
            tri_coords(p, q, r, px, bc);
            if (bc[0] >= 0 && bc[1] >= 0 && bc[2] >= 0) {
                const float fr = (float)(bc[0] * t[0].c[0] + bc[1] * t[1].c[0] + bc[2] * t[2].c[0]);
                const float fg = (float)(bc[0] * t[0].c[1] + bc[1] * t[1].c[1] + bc[2] * t[2].c[1]);
                const float fb = (float)(bc[0] * t[0].c[2] + bc[1] * t[1].c[2] + bc[2] * t[2].c[2]);
                const int row = res - j, col = i;
                if (row < 0 || row >= res || col < 0 || col >= res) continue;    /* OOB in the reference */
                uint8_t *o = bgra + 4 * ((size_t)row * res + col);
                o[0] = (uint8_t)fb; o[1] = (uint8_t)fg; o[2] = (uint8_t)fr; o[3] = 255;
            }
        }
    }
}

#define PTO_MAX_NB 96      /* 3 * PT_MAX_K */

/* The whole face loop (src/pointsTransfer.cpp:465-585).  idx[n_vertices * k]: neighbour lists of
 * the mesh vertices (ascending (d2, index), -1 padded) over the cloud `pts`.  bgra: res*res*4,
 * zero-initialised by the caller.  stats[0] += sub-triangles drawn, stats[1] += inside points. */
int pto_texture_faces(const pto_point *pts, const pto_point *vertices, const int32_t *idx, int k,
                      const int32_t *faces, int64_t n_faces, int res, uint8_t *bgra, int64_t *stats)
{
    if (k < 1 || 3 * k > PTO_MAX_NB) return 1;
    for (int64_t f = 0; f < n_faces; ++f) {
        const pto_point *tv[3] = {&vertices[faces[3 * f]], &vertices[faces[3 * f + 1]], &vertices[faces[3 * f + 2]]};
        int nb[PTO_MAX_NB], n_nb = 0;
        for (int c = 0; c < 3; ++c)
            for (int j = 0; j < k; ++j) {
                const int id = idx[(size_t)faces[3 * f + c] * k + j];
                if (id < 0) continue;
                int seen = 0;
                for (int e = 0; e < n_nb; ++e) seen |= nb[e] == id;
                if (!seen) nb[n_nb++] = id;
            }
        double n[3], b1[3], b2[3];
        plane_frame(tv[0]->ver, tv[1]->ver, tv[2]->ver, n, b1, b2);
        v2 P[3 + PTO_MAX_NB];
        double BC[PTO_MAX_NB][3];
        int who[PTO_MAX_NB], np = 3;
        for (int c = 0; c < 3; ++c) P[c] = to_2d(tv[c]->ver, tv[0]->ver, n, b1, b2);
        for (int e = 0; e < n_nb; ++e) {
            const v2 x = to_2d(pts[nb[e]].ver, tv[0]->ver, n, b1, b2);
            double bc[3];
            tri_coords(P[0], P[1], P[2], x, bc);
            if (bc[0] >= 0 && bc[1] >= 0 && bc[2] >= 0) {
                P[np] = x;
                memcpy(BC[np - 3], bc, sizeof bc);
                who[np - 3] = nb[e];
                ++np;
            }
        }
        tex_vertex V[3 + PTO_MAX_NB];
        for (int c = 0; c < 3; ++c) {
            V[c].u = tv[c]->U; V[c].v = tv[c]->V;
            for (int a = 0; a < 3; ++a) V[c].c[a] = tv[c]->color[a];
        }
        if (np == 3) {
            draw_triangle(V, res, bgra);
            if (stats) stats[0] += 1;
            continue;
        }
        for (int e = 3; e < np; ++e) {
            const double *bc = BC[e - 3];
            V[e].u = bc[0] * tv[0]->U + bc[1] * tv[1]->U + bc[2] * tv[2]->U;      /* :572 */
            V[e].v = bc[0] * tv[0]->V + bc[1] * tv[1]->V + bc[2] * tv[2]->V;      /* :573 */
            for (int a = 0; a < 3; ++a) V[e].c[a] = pts[who[e - 3]].color[a];
        }
        if (stats) stats[1] += np - 3;
        for (int a = 0; a < np; ++a)
            for (int b = a + 1; b < np; ++b)
                for (int c = b + 1; c < np; ++c) {
                    const double o = area2(P[a], P[b], P[c]);
                    if (o == 0.0 || o != o) continue;
                    int ok = 1;
                    for (int d = 0; d < np && ok; ++d) {
                        if (d == a || d == b || d == c) continue;
                        /* The in-circle determinant is alternating in its 4 points: evaluate it
                         * once on the index-sorted quadruple and give it the sign of the
                         * permutation (d moves from the last place to its sorted place by `sw`
                         * adjacent swaps), so the four predicates of a quadruple can never
                         * contradict each other.  d is strictly inside the circle through
                         * (a, b, c) iff incircle(a, b, c, d) * orientation(a, b, c) > 0. */
                        int q4[4] = {a, b, c, d}, sw = 0;
                        for (int i = 3; i > 0 && q4[i] < q4[i - 1]; --i) { int t = q4[i]; q4[i] = q4[i - 1]; q4[i - 1] = t; ++sw; }
                        double s = incircle(P[q4[0]], P[q4[1]], P[q4[2]], P[q4[3]]);
                        if (sw & 1) s = -s;
                        if (o < 0) s = -s;
                        if (s > 0) ok = 0;
                    }
                    if (!ok) continue;
                    const tex_vertex T[3] = {V[a], V[b], V[c]};
                    draw_triangle(T, res, bgra);
                    if (stats) stats[0] += 1;
                }
    }
    return 0;
}

/* src/pointsTransfer.cpp:593-611: dilate 25x25 (border ignored), edges = dilated & ~alpha,
 * padded = saturate(texture + edges).  in / out: res*res*4 BGRA. */
int pto_texture_pad(const uint8_t *in, int res, uint8_t *out)
{
    const int R = 12;
    uint8_t *tmp = (uint8_t *)malloc((size_t)res * res * 4);
    if (!tmp) return 1;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < res; ++y)
        for (int x = 0; x < res; ++x)
            for (int c = 0; c < 4; ++c) {
                uint8_t m = 0;
                for (int dx = -R; dx <= R; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= res) continue;
                    const uint8_t v = in[4 * ((size_t)y * res + xx) + c];
                    if (v > m) m = v;
                }
                tmp[4 * ((size_t)y * res + x) + c] = m;
            }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < res; ++y)
        for (int x = 0; x < res; ++x) {
            const uint8_t alpha = in[4 * ((size_t)y * res + x) + 3];
            for (int c = 0; c < 4; ++c) {
                uint8_t m = 0;
                for (int dy = -R; dy <= R; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= res) continue;
                    const uint8_t v = tmp[4 * ((size_t)yy * res + x) + c];
                    if (v > m) m = v;
                }
                const int edge = m & (uint8_t)~alpha;
                const int s = in[4 * ((size_t)y * res + x) + c] + edge;
                out[4 * ((size_t)y * res + x) + c] = (uint8_t)(s > 255 ? 255 : s);
            }
        }
    free(tmp);
    return 0;
}
