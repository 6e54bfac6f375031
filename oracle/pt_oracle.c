/*
 * pt_oracle.c -- CPU oracle (test infrastructure, see pt_oracle.h header).
 * Metric / box bound / record layout: pinned bit for bit on the reference's own headers
 * (oracle/ref_shim.cpp -> oracle/_ref, tests/test_oracle.py).  Neighbour SEARCH: parity
 * unpinned (no reference golden vectors exist; CGAL absent).
 *
 * Build: see oracle/Makefile.  Compiled like the reference's Release build
 * (src/CMakeLists.txt:7-9: -O3, no -march, no -ffast-math), plus
 * -ffp-contract=off so no FMA is ever formed (x86-64 baseline has none).
 */
#include "pt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

_Static_assert(sizeof(pto_point) == 80, "Point must be 80 bytes (src/Point.h:1-6)");
_Static_assert(offsetof(pto_point, normal) == 24, "normal @24");
_Static_assert(offsetof(pto_point, color) == 48, "color @48");
_Static_assert(offsetof(pto_point, U) == 64, "U @64");
_Static_assert(offsetof(pto_point, V) == 72, "V @72");

int pto_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- metric and bounds: src/Distance.h ---------------------------------- */

/* src/Distance.h:6-11 */
double pto_transformed_distance(const pto_point *p1, const pto_point *p2)
{
    double distx = p1->ver[0] - p2->ver[0];
    double disty = p1->ver[1] - p2->ver[1];
    double distz = p1->ver[2] - p2->ver[2];
    return distx * distx + disty * disty + distz * distz;
}

static inline double dist3(const double *a, const double *b)
{
    double distx = a[0] - b[0];
    double disty = a[1] - b[1];
    double distz = a[2] - b[2];
    return distx * distx + disty * disty + distz * distz;
}

/* src/Distance.h:27-57 (the 3-argument form used by orthogonal search) */
double pto_min_distance_to_rectangle(const pto_point *p, const double lo[3],
                                     const double hi[3], double dists[3])
{
    double distance = 0.0;
    for (int a = 0; a < 3; ++a) {
        double h = p->ver[a];
        if (h < lo[a]) {
            dists[a] = lo[a] - h;
            distance += dists[a] * dists[a];
        }
        if (h > hi[a]) {
            dists[a] = h - hi[a];
            distance += dists[a] * dists[a];
        }
    }
    return distance;
}

/* src/Distance.h:92-95 */
double pto_new_distance(double dist, double old_off, double new_off)
{
    return dist + new_off * new_off - old_off * old_off;
}

/* src/Distance.h:97 */
double pto_transformed_radius(double d) { return d * d; }

static double radius_to_r2(double radius)
{
    if (!(radius >= 0.0) || isinf(radius)) return INFINITY;
    return pto_transformed_radius(radius);
}

/* ---- ordering key (d2, idx) --------------------------------------------- */

typedef struct { double d2; int32_t idx; } cand_t;

static inline int cand_less(double d2a, int32_t ia, double d2b, int32_t ib)
{
    return d2a < d2b || (d2a == d2b && ia < ib);
}

/* sorted insertion into an ascending list of at most k entries */
static inline void list_insert(cand_t *list, int *cnt, int k, double d2, int32_t idx)
{
    int c = *cnt;
    if (c == k) {
        if (!cand_less(d2, idx, list[k - 1].d2, list[k - 1].idx)) return;
        c = k - 1;
    }
    int j = c;
    while (j > 0 && cand_less(d2, idx, list[j - 1].d2, list[j - 1].idx)) {
        list[j] = list[j - 1];
        --j;
    }
    list[j].d2 = d2;
    list[j].idx = idx;
    *cnt = c + 1;
}

static void emit(const cand_t *list, int cnt, int k, int32_t *idx_out, double *d2_out)
{
    for (int j = 0; j < k; ++j) {
        idx_out[j] = j < cnt ? list[j].idx : -1;
        if (d2_out) d2_out[j] = j < cnt ? list[j].d2 : INFINITY;
    }
}

/* ---- brute force -------------------------------------------------------- */

int pto_knn_bruteforce(const pto_point *pts, int64_t n, const pto_point *queries,
                       int64_t m, int k, double radius, int32_t *idx_out,
                       double *d2_out, int nthreads)
{
    if (k <= 0) return 1;
    const double r2 = radius_to_r2(radius);
    double *xyz = (double *)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(double));
    if (!xyz) return 2;
    for (int64_t i = 0; i < n; ++i) {
        xyz[3 * i + 0] = pts[i].ver[0];
        xyz[3 * i + 1] = pts[i].ver[1];
        xyz[3 * i + 2] = pts[i].ver[2];
    }
    if (nthreads <= 0) nthreads = pto_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        cand_t *list = (cand_t *)malloc((size_t)k * sizeof(cand_t));
#pragma omp for schedule(dynamic, 4)
        for (int64_t q = 0; q < m; ++q) {
            const double *qv = queries[q].ver;
            int cnt = 0;
            double worst = r2; /* candidate must have d2 <= worst to matter */
            for (int64_t i = 0; i < n; ++i) {
                double d2 = dist3(qv, xyz + 3 * i);
                if (d2 > worst) continue;
                list_insert(list, &cnt, k, d2, (int32_t)i);
                if (cnt == k && list[k - 1].d2 < worst) worst = list[k - 1].d2;
            }
            emit(list, cnt, k, idx_out + q * k, d2_out ? d2_out + q * k : NULL);
        }
        free(list);
    }
    free(xyz);
    return 0;
}

/* ---- blend (frozen definition; the reference has none) ----
 *
 * For one sample with neighbours j = 0..cnt-1 in ascending (d2, index) order:
 *   weights  w_j = 1/d2_j;  if d2_0 == 0 (exact hit): w_j = (d2_j == 0) ? 1 : 0;
 *            if the weight sum overflows (not finite): w_0 = 1, others 0.
 *   sums     W = sum w_j, C_c = sum w_j*colour_jc, N_a = sum w_j*(double)(float)normal_ja,
 *            every sum accumulated sequentially in neighbour order, fp64, no FMA.
 *   colour   trunc(C_c / W) clamped to 0..255 (truncation as src/pointsTransfer.cpp:100-102),
 *            alpha 255 (:103); input colours are clamped to 0..255 first.
 *   normal   N / sqrt((Nx*Nx + Ny*Ny) + Nz*Nz) rounded to fp32; zero if the length is 0.
 *   cnt == 0 rgba = 0,0,0,0 and normal = 0,0,0.
 */
#define PTO_MAX_K 32

int pto_blend(const pto_point *pts, int64_t n, int64_t m, int k,
              const int32_t *idx, const double *d2, uint8_t *rgba_out,
              float *normal_out)
{
    if (k <= 0 || k > PTO_MAX_K) return 1;
    int bad_index = 0;
#pragma omp parallel for schedule(static) reduction(| : bad_index)
    for (int64_t q = 0; q < m; ++q) {
        const int32_t *qi = idx + q * k;
        const double *qd = d2 + q * k;
        int cnt = 0;
        while (cnt < k && qi[cnt] >= 0) ++cnt;
        uint8_t *rgba = rgba_out + 4 * q;
        float *nrm = normal_out + 3 * q;
        if (cnt == 0) {
            rgba[0] = rgba[1] = rgba[2] = rgba[3] = 0;
            nrm[0] = nrm[1] = nrm[2] = 0.0f;
            continue;
        }
        /* 0: 1/d2, 1: exact hits only (d2 == 0), 2: nearest only (weights overflowed) */
        int mode = (qd[0] == 0.0) ? 1 : 0;
        double s[7];
        for (int pass = 0; pass < 2; ++pass) {
            for (int a = 0; a < 7; ++a) s[a] = 0.0;
            for (int j = 0; j < cnt; ++j) {
                double w;
                if (mode == 0) w = 1.0 / qd[j];
                else if (mode == 1) w = (qd[j] == 0.0) ? 1.0 : 0.0;
                else w = (j == 0) ? 1.0 : 0.0;
                if (qi[j] >= n) bad_index = 1;
                const pto_point *p = &pts[qi[j] >= n ? 0 : qi[j]];
                s[0] = s[0] + w;
                for (int c = 0; c < 3; ++c) {
                    int col = p->color[c];
                    col = col < 0 ? 0 : (col > 255 ? 255 : col);
                    s[1 + c] = s[1 + c] + w * (double)col;
                    s[4 + c] = s[4 + c] + w * (double)(float)p->normal[c];
                }
            }
            if (s[0] > 0.0 && s[0] < INFINITY) break;
            mode = 2;
        }
        for (int c = 0; c < 3; ++c) {
            double v = s[1 + c] / s[0];
            int iv = (int)v; /* truncation, src/pointsTransfer.cpp:100-102 */
            rgba[c] = (uint8_t)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
        }
        rgba[3] = 255; /* src/pointsTransfer.cpp:103 */
        double len = sqrt(s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
        if (len > 0.0 && len < INFINITY) {
            nrm[0] = (float)(s[4] / len);
            nrm[1] = (float)(s[5] / len);
            nrm[2] = (float)(s[6] / len);
        } else {
            nrm[0] = nrm[1] = nrm[2] = 0.0f;
        }
    }
    return bad_index ? 3 : 0;
}

/* ---- kd-tree: CGAL Kd_tree<Sliding_midpoint, bucket 10> (recalled) ------- */

typedef struct {
    /* internal: cut_dim 0..2, children indices; leaf: cut_dim = -1 */
    int32_t cut_dim;
    int32_t lower, upper;        /* child node ids (internal) */
    int64_t begin, end;          /* range in ptrs[] (leaf) */
    double  cut_val;
    double  lower_low, lower_high, upper_low, upper_high;
} kd_node;

struct pto_kdtree {
    const pto_point  *base;   /* caller's array (index = ptr - base) */
    pto_point        *pts;    /* CGAL copies the points into the tree (:259) */
    const pto_point **ptrs;   /* partitioned pointers, leaf-contiguous */
    kd_node          *nodes;
    int64_t           n, n_nodes, cap_nodes;
    int               bucket;
    double            bb_lo[3], bb_hi[3];
};

static void tight_box(const pto_point **p, int64_t b, int64_t e, double lo[3], double hi[3])
{
    for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
    for (int64_t i = b; i < e; ++i)
        for (int a = 0; a < 3; ++a) {
            double v = p[i]->ver[a];
            if (v < lo[a]) lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
}

/* Nodes come from one array reserved up front (a tree over n points has fewer than 2n nodes;
 * untouched pages cost nothing), so subtrees can be built by concurrent OpenMP tasks: the only
 * shared state is this counter. */
static int32_t new_node(pto_kdtree *t)
{
    int64_t id;
#pragma omp atomic capture
    id = t->n_nodes++;
    return (int32_t)id;
}

/* Builds the subtree over ptrs[b,e) whose loose box is (blo,bhi) and tight box
 * (tlo,thi).  Sliding midpoint (CGAL Splitters.h, recalled): cut the longest
 * side of the loose box at its midpoint; if the tight box is degenerate there,
 * use the longest tight side; slide the cut onto the tight box if all points
 * fall on one side. */
#define PTO_TASK_MIN 65536
static int32_t build_rec(pto_kdtree *t, int64_t b, int64_t e, const double blo[3],
                         const double bhi[3], const double tlo[3], const double thi[3])
{
    int32_t id = new_node(t);
    if (e - b <= t->bucket) {
        kd_node *nd = &t->nodes[id];
        nd->cut_dim = -1; nd->begin = b; nd->end = e; nd->lower = nd->upper = -1;
        return id;
    }
    int cd = 0;
    for (int a = 1; a < 3; ++a) if (bhi[a] - blo[a] > bhi[cd] - blo[cd]) cd = a;
    double cut;
    if (tlo[cd] != thi[cd]) {
        cut = (bhi[cd] + blo[cd]) / 2.0;
    } else {
        cd = 0;
        for (int a = 1; a < 3; ++a) if (thi[a] - tlo[a] > thi[cd] - tlo[cd]) cd = a;
        cut = (thi[cd] + tlo[cd]) / 2.0;
    }
    if (thi[cd] <= cut) cut = thi[cd];
    if (tlo[cd] >= cut) cut = tlo[cd];
    /* partition: coord < cut -> lower; sliding keeps both sides non-empty */
    const pto_point **p = t->ptrs;
    int64_t i = b, j = e;
    while (i < j) {
        if (p[i]->ver[cd] < cut) ++i;
        else { --j; const pto_point *tmp = p[i]; p[i] = p[j]; p[j] = tmp; }
    }
    int64_t mid = i;
    if (mid == b) {
        /* all points >= cut (cut == tight low): move one minimal point down */
        int64_t best = b;
        for (int64_t q = b + 1; q < e; ++q) if (p[q]->ver[cd] < p[best]->ver[cd]) best = q;
        const pto_point *tmp = p[b]; p[b] = p[best]; p[best] = tmp;
        mid = b + 1;
    } else if (mid == e) {
        int64_t best = b;
        for (int64_t q = b + 1; q < e; ++q) if (p[q]->ver[cd] > p[best]->ver[cd]) best = q;
        const pto_point *tmp = p[e - 1]; p[e - 1] = p[best]; p[best] = tmp;
        mid = e - 1;
    }
    double llo[3], lhi[3], ulo[3], uhi[3], lblo[3], lbhi[3], ublo[3], ubhi[3];
    tight_box(p, b, mid, llo, lhi);
    tight_box(p, mid, e, ulo, uhi);
    memcpy(lblo, blo, sizeof lblo); memcpy(lbhi, bhi, sizeof lbhi);
    memcpy(ublo, blo, sizeof ublo); memcpy(ubhi, bhi, sizeof ubhi);
    lbhi[cd] = cut; ublo[cd] = cut;
    {
        kd_node *nd = &t->nodes[id];
        nd->cut_dim = cd; nd->cut_val = cut; nd->begin = b; nd->end = e;
        nd->lower_low = llo[cd]; nd->lower_high = lhi[cd];
        nd->upper_low = ulo[cd]; nd->upper_high = uhi[cd];
    }
    int32_t lo_id, up_id;
    if (e - b >= PTO_TASK_MIN) {      /* big subtrees: one task each (disjoint ranges of ptrs[]) */
#pragma omp task shared(lo_id) firstprivate(t, b, mid, lblo, lbhi, llo, lhi)
        lo_id = build_rec(t, b, mid, lblo, lbhi, llo, lhi);
        up_id = build_rec(t, mid, e, ublo, ubhi, ulo, uhi);
#pragma omp taskwait
    } else {
        lo_id = build_rec(t, b, mid, lblo, lbhi, llo, lhi);
        up_id = build_rec(t, mid, e, ublo, ubhi, ulo, uhi);
    }
    t->nodes[id].lower = lo_id;
    t->nodes[id].upper = up_id;
    return id;
}

pto_kdtree *pto_kdtree_build(const pto_point *pts, int64_t n, int bucket_size)
{
    pto_kdtree *t = (pto_kdtree *)calloc(1, sizeof *t);
    if (!t) return NULL;
    t->base = pts; t->n = n; t->bucket = bucket_size > 0 ? bucket_size : 10;
    t->pts = (pto_point *)malloc((size_t)(n > 0 ? n : 1) * sizeof(pto_point));
    t->ptrs = (const pto_point **)malloc((size_t)(n > 0 ? n : 1) * sizeof(*t->ptrs));
    if (!t->pts || !t->ptrs) { pto_kdtree_free(t); return NULL; }
    t->cap_nodes = 2 * (n > 0 ? n : 1) + 16;
    t->nodes = (kd_node *)malloc((size_t)t->cap_nodes * sizeof(kd_node));
    if (!t->nodes) { pto_kdtree_free(t); return NULL; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) { t->pts[i] = pts[i]; t->ptrs[i] = &t->pts[i]; }
    if (n > 0) {
        tight_box(t->ptrs, 0, n, t->bb_lo, t->bb_hi);
#pragma omp parallel
#pragma omp single
        build_rec(t, 0, n, t->bb_lo, t->bb_hi, t->bb_lo, t->bb_hi);
    }
    {   /* give the untouched tail of the reservation back */
        kd_node *shrunk = (kd_node *)realloc(t->nodes, (size_t)(t->n_nodes > 0 ? t->n_nodes : 1) * sizeof(kd_node));
        if (shrunk) { t->nodes = shrunk; t->cap_nodes = t->n_nodes; }
    }
    return t;
}

void pto_kdtree_free(pto_kdtree *t)
{
    if (!t) return;
    free(t->pts); free(t->ptrs); free(t->nodes); free(t);
}

int64_t pto_kdtree_node_count(const pto_kdtree *t) { return t->n_nodes; }

/* bounded max-heap on (d2, idx): top = current worst */
typedef struct {
    cand_t *h; int cnt, k; int exact; double r2;
} kq_t;

static inline int heap_before(const cand_t *a, const cand_t *b)
{ /* max-heap order: a is "larger" than b */
    return cand_less(b->d2, b->idx, a->d2, a->idx);
}

static void heap_sift_down(cand_t *h, int n, int i)
{
    for (;;) {
        int l = 2 * i + 1, r = l + 1, big = i;
        if (l < n && heap_before(&h[l], &h[big])) big = l;
        if (r < n && heap_before(&h[r], &h[big])) big = r;
        if (big == i) return;
        cand_t tmp = h[i]; h[i] = h[big]; h[big] = tmp; i = big;
    }
}

static void heap_sift_up(cand_t *h, int i)
{
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!heap_before(&h[i], &h[p])) return;
        cand_t tmp = h[i]; h[i] = h[p]; h[p] = tmp; i = p;
    }
}

static inline void kq_offer(kq_t *q, double d2, int32_t idx)
{
    if (d2 > q->r2) return;
    if (q->cnt < q->k) {
        q->h[q->cnt].d2 = d2; q->h[q->cnt].idx = idx;
        heap_sift_up(q->h, q->cnt++);
        return;
    }
    int better = q->exact ? cand_less(d2, idx, q->h[0].d2, q->h[0].idx)
                          : (d2 < q->h[0].d2); /* CGAL: strict '<' */
    if (!better) return;
    q->h[0].d2 = d2; q->h[0].idx = idx;
    heap_sift_down(q->h, q->cnt, 0);
}

static inline int kq_branch(const kq_t *q, double rd)
{
    if (rd > q->r2) return 0;
    if (q->cnt < q->k) return 1;
    /* CGAL branch_nearest: rd * (1+eps)^2 < top, eps = 0 (Distance.h:97) */
    return q->exact ? (rd <= q->h[0].d2) : (rd * pto_transformed_radius(1.0) < q->h[0].d2);
}

typedef struct {
    const pto_kdtree *t; const double *qv; kq_t *kq; double dists[3];
} search_t;

static void search_rec(search_t *s, int32_t id, double rd)
{
    const kd_node *nd = &s->t->nodes[id];
    if (nd->cut_dim < 0) {
        const pto_point **p = s->t->ptrs;
        for (int64_t i = nd->begin; i < nd->end; ++i) {
            double d2 = dist3(s->qv, p[i]->ver); /* Distance.h:6-11 */
            kq_offer(s->kq, d2, (int32_t)(p[i] - s->t->pts));
        }
        return;
    }
    int cd = nd->cut_dim;
    double val = s->qv[cd];
    double diff1 = val - nd->upper_low;
    double diff2 = val - nd->lower_high;
    int32_t best, other; double new_off;
    if (diff1 + diff2 < 0) { new_off = diff1; best = nd->lower; other = nd->upper; }
    else                   { new_off = diff2; best = nd->upper; other = nd->lower; }
    search_rec(s, best, rd);
    double dst = s->dists[cd];
    double new_rd = pto_new_distance(rd, dst, new_off); /* Distance.h:92-95 */
    s->dists[cd] = new_off;
    if (kq_branch(s->kq, new_rd)) search_rec(s, other, new_rd);
    s->dists[cd] = dst;
}

static int heap_cmp_asc(const void *a, const void *b)
{
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (cand_less(x->d2, x->idx, y->d2, y->idx)) return -1;
    if (cand_less(y->d2, y->idx, x->d2, x->idx)) return 1;
    return 0;
}

static int knn_one(const pto_kdtree *t, const pto_point *query, int k, double r2,
                   int exact, cand_t *heap)
{
    kq_t kq = { heap, 0, k, exact, r2 };
    if (t->n == 0) return 0;
    search_t s; s.t = t; s.qv = query->ver; s.kq = &kq;
    s.dists[0] = s.dists[1] = s.dists[2] = 0.0;
    double rd = pto_min_distance_to_rectangle(query, t->bb_lo, t->bb_hi, s.dists);
    search_rec(&s, 0, rd);
    qsort(heap, (size_t)kq.cnt, sizeof(cand_t), heap_cmp_asc);
    return kq.cnt;
}

int pto_kdtree_knn(const pto_kdtree *t, const pto_point *queries, int64_t m,
                   int k, double radius, int exact_ties, int32_t *idx_out,
                   double *d2_out, int nthreads)
{
    if (k <= 0) return 1;
    const double r2 = radius_to_r2(radius);
    if (nthreads <= 0) nthreads = pto_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        cand_t *heap = (cand_t *)malloc((size_t)k * sizeof(cand_t));
#pragma omp for schedule(dynamic, 64)
        for (int64_t q = 0; q < m; ++q) {
            int cnt = knn_one(t, &queries[q], k, r2, exact_ties, heap);
            emit(heap, cnt, k, idx_out + q * k, d2_out ? d2_out + q * k : NULL);
        }
        free(heap);
    }
    return 0;
}

/* src/pointsTransfer.cpp:465-479 -- 3 K-NN searches per face, one per corner;
 * the results are consumed (copied, as the reference copies 80-byte Points into
 * its std::set at :477) so the work cannot be optimised away. */
int64_t pto_reference_face_loop(const pto_kdtree *t, const pto_point *vertices,
                                const int32_t *faces, int64_t face_count, int k,
                                int nthreads)
{
    int64_t produced = 0;
    if (nthreads <= 0) nthreads = pto_max_threads();
#pragma omp parallel num_threads(nthreads) reduction(+ : produced)
    {
        cand_t *heap = (cand_t *)malloc((size_t)k * sizeof(cand_t));
        pto_point *neighbors = (pto_point *)malloc((size_t)3 * k * sizeof(pto_point));
#pragma omp for schedule(dynamic, 64)
        for (int64_t j = 0; j < face_count; ++j) {
            int nn = 0;
            for (int i = 0; i < 3; ++i) {
                const pto_point *corner = &vertices[faces[3 * j + i]];
                int cnt = knn_one(t, corner, k, INFINITY, 0, heap);
                for (int c = 0; c < cnt; ++c) neighbors[nn++] = t->pts[heap[c].idx];
            }
            produced += nn;
            if (nn && neighbors[nn - 1].ver[0] != neighbors[nn - 1].ver[0]) produced = -1;
        }
        free(heap); free(neighbors);
    }
    return produced;
}
