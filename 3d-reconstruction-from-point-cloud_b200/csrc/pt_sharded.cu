// pt_sharded.cu -- a cloud sharded over several GPUs of one box behind the C ABI, host side in
// C++ (north_star: "host code stays C++"; "each GPU indexes a spatial slab of the cloud plus a
// halo sized to the search radius; queries are ... routed by slab").  One process, one host
// thread per slab; built only on the single-index entry points of this library:
//
//   build   : x-cuts at point-count quantiles -> per slab its own points plus the GHOST ZONE
//             (the other slabs' points within `halo` of the slab's box), global ids ascending,
//             one pt_index per slab on its device (pt_index_build)
//   transfer: samples routed to the slab whose x-range holds them; every slab answers its
//             samples with pt_transfer_slab -- k-NN + blend + the ghost-zone check fused per
//             pipeline chunk -- and the results land in the caller's arrays in sample order.
//             A step is final iff no sample's k-th-neighbour ball can leave its slab's ghost
//             zone towards another slab (exactness argument: DESIGN.md section 6).  If some
//             slab reports needs_exchange the halo is widened (x4, at least the bounding-box estimate), the slabs are rebuilt and the
//             call is repeated, so the result is always exact; `rebuilds` counts that.
//
// The reference has no multi-GPU path (its query loop is src/pointsTransfer.cpp:465-479); the
// torch.distributed / NCCL form of the same sharding (one process per GPU, unsorted samples
// routed over all_to_all) is sharded.py.  Devices may repeat in the device list: several slabs
// then share one GPU, which is how the single-GPU test box exercises this file.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "points_transfer.h"

namespace {

struct Rec80 {
    double ver[3];
    double normal[3];
    int    color[3];
    int    pad;
    double U, V;
};
static_assert(sizeof(Rec80) == PT_POINT_STRIDE, "Point must be 80 bytes");

struct Slab {
    pt_index *index = nullptr;
    int device = 0;
    size_t n_own = 0, n_ghost = 0;
    double box[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};   // own points
};

}  // namespace

struct pt_sharded {
    std::vector<Slab> slabs;
    std::vector<double> cuts;         // n_slabs + 1, first -inf, last +inf
    std::vector<double> boxes;        // n_slabs x 6, the slabs' OWN boxes
    const Rec80 *points = nullptr;    // the caller's cloud (must outlive the handle)
    size_t n = 0;
    double halo = 0.0;
    double halo_auto = 0.0;           // the bounding-box estimate (4x the expected k-th neighbour distance)
    int coord_mode = PT_COORD_AUTO;
    int rebuilds = 0;
};

namespace {

void free_slabs(pt_sharded *S)
{
    for (auto &s : S->slabs) {
        if (s.index) pt_index_free(s.index);
        s.index = nullptr;
    }
}

// (re)builds every slab index for the current halo
int build_slabs(pt_sharded *S)
{
    const int R = (int)S->slabs.size();
    free_slabs(S);
    const Rec80 *P = S->points;
    const size_t n = S->n;
    // owner of every point, own boxes
    std::vector<int> owner(n);
    for (auto &s : S->slabs) {
        s.n_own = s.n_ghost = 0;
        for (int a = 0; a < 3; ++a) { s.box[a] = INFINITY; s.box[3 + a] = -INFINITY; }
    }
    for (size_t i = 0; i < n; ++i) {
        const double x = P[i].ver[0];
        const int r = (int)(std::upper_bound(S->cuts.begin() + 1, S->cuts.end() - 1, x) - (S->cuts.begin() + 1));
        owner[i] = r;
        Slab &s = S->slabs[r];
        ++s.n_own;
        for (int a = 0; a < 3; ++a) {
            s.box[a] = std::min(s.box[a], P[i].ver[a]);
            s.box[3 + a] = std::max(s.box[3 + a], P[i].ver[a]);
        }
    }
    S->boxes.assign((size_t)R * 6, 0.0);
    for (int r = 0; r < R; ++r) memcpy(&S->boxes[(size_t)r * 6], S->slabs[r].box, sizeof(double) * 6);
    std::vector<int> status(R, PT_OK);
    std::vector<std::thread> th;
    const double h2 = S->halo * S->halo * (1.0 + 1e-9);
    auto work = [&](int r) {
          try {
            Slab &s = S->slabs[r];
            std::vector<Rec80> rec;
            std::vector<int32_t> ids;
            rec.reserve(s.n_own + s.n_own / 16 + 1024);
            ids.reserve(s.n_own + s.n_own / 16 + 1024);
            const double *b = s.box;
            for (size_t i = 0; i < n; ++i) {          // index order: global ids ascending
                bool take = owner[i] == r;
                if (!take && s.n_own) {
                    const double ex = std::max(std::max(b[0] - P[i].ver[0], P[i].ver[0] - b[3]), 0.0);
                    const double ey = std::max(std::max(b[1] - P[i].ver[1], P[i].ver[1] - b[4]), 0.0);
                    const double ez = std::max(std::max(b[2] - P[i].ver[2], P[i].ver[2] - b[5]), 0.0);
                    take = ex * ex + ey * ey + ez * ez <= h2;
                }
                if (take) { rec.push_back(P[i]); ids.push_back((int32_t)i); }
            }
            s.n_ghost = rec.size() - s.n_own;
            pt_build_opts o;
            memset(&o, 0, sizeof o);
            o.device = s.device;
            o.coord_mode = S->coord_mode;
            o.ids = ids.data();
            status[r] = pt_index_build(rec.data(), rec.size(), &o, &s.index);
          } catch (...) { status[r] = PT_ERR_OUT_OF_MEMORY; }     // nothing is thrown across the ABI
    };
    for (int r = 0; r < R; ++r) {
        try { th.emplace_back(work, r); } catch (...) { work(r); }      // no thread to be had: inline
    }
    for (auto &t : th) t.join();
    for (int r = 0; r < R; ++r)
        if (status[r] != PT_OK) { free_slabs(S); return status[r]; }
    return PT_OK;
}

int sharded_query(pt_sharded *S, const void *queries, size_t m, int k, double radius, int32_t *idx_out,
                  double *d2_out, uint8_t *rgba_out, float *normal_out)
{
    if (!S || (!queries && m)) return PT_ERR_INVALID_ARG;
    if (k < 1 || k > PT_MAX_K) return PT_ERR_UNSUPPORTED;
    if (m == 0) return PT_OK;
    const int R = (int)S->slabs.size();
    const Rec80 *Q = (const Rec80 *)queries;
    // route: the slab whose x-range holds the sample
    std::vector<std::vector<uint32_t>> mine(R);
    for (size_t i = 0; i < m; ++i) {
        const double x = Q[i].ver[0];
        const int r = (int)(std::upper_bound(S->cuts.begin() + 1, S->cuts.end() - 1, x) - (S->cuts.begin() + 1));
        mine[r].push_back((uint32_t)i);
    }
    for (int attempt = 0; attempt < 8; ++attempt) {
        std::vector<int> status(R, PT_OK), needs(R, 0);
        std::vector<std::thread> th;
        auto work = [&](int r) {
              try {
                const std::vector<uint32_t> &sel = mine[r];
                const size_t mr = sel.size();
                if (mr == 0) return;
                std::vector<Rec80> q(mr);
                for (size_t j = 0; j < mr; ++j) q[j] = Q[sel[j]];
                std::vector<int32_t> idx(idx_out ? mr * k : 0);
                std::vector<double> d2(d2_out ? mr * k : 0);
                std::vector<uint8_t> rgba(rgba_out ? mr * 4 : 0);
                std::vector<float> nrm(normal_out ? mr * 3 : 0);
                int need = 0;
                int rc;
                if (S->slabs[r].index == nullptr) { status[r] = PT_ERR_INVALID_ARG; return; }
                if (rgba_out || normal_out)
                    rc = pt_transfer_slab(S->slabs[r].index, q.data(), 0, mr, k, radius, S->boxes.data(), R, r, S->halo,
                                          idx_out ? idx.data() : nullptr, d2_out ? d2.data() : nullptr,
                                          rgba_out ? rgba.data() : nullptr, normal_out ? nrm.data() : nullptr, &need);
                else {       // k-NN only: the slab entry needs a blend output to run; use a scratch one
                    rgba.resize(mr * 4);
                    rc = pt_transfer_slab(S->slabs[r].index, q.data(), 0, mr, k, radius, S->boxes.data(), R, r, S->halo,
                                          idx_out ? idx.data() : nullptr, d2_out ? d2.data() : nullptr, rgba.data(),
                                          nullptr, &need);
                }
                status[r] = rc;
                needs[r] = need;
                if (rc != PT_OK || need) return;
                for (size_t j = 0; j < mr; ++j) {          // results back into sample order
                    const size_t d = sel[j];
                    if (idx_out) memcpy(idx_out + d * k, idx.data() + j * k, sizeof(int32_t) * k);
                    if (d2_out) memcpy(d2_out + d * k, d2.data() + j * k, sizeof(double) * k);
                    if (rgba_out) memcpy(rgba_out + d * 4, rgba.data() + j * 4, 4);
                    if (normal_out) memcpy(normal_out + d * 3, nrm.data() + j * 3, sizeof(float) * 3);
                }
              } catch (...) { status[r] = PT_ERR_OUT_OF_MEMORY; }
        };
        for (int r = 0; r < R; ++r) {
            try { th.emplace_back(work, r); } catch (...) { work(r); }
        }
        for (auto &t : th) t.join();
        bool again = false;
        for (int r = 0; r < R; ++r) {
            if (status[r] != PT_OK) return status[r];
            again |= needs[r] != 0;
        }
        if (!again) return PT_OK;
        // some k-th-neighbour ball may leave a ghost zone: widen it and answer the call again
        S->halo = std::max(4.0 * S->halo, S->halo_auto);
        ++S->rebuilds;
        const int rc = build_slabs(S);
        if (rc != PT_OK) return rc;
    }
    return PT_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" {

int pt_sharded_build(const void *points, size_t n, const pt_sharded_opts *opts, pt_sharded **out)
{
    if (!out || !opts || (!points && n) || n >= 0x7fffffffull || opts->n_devices < 1 || opts->n_devices > 64)
        return PT_ERR_INVALID_ARG;
    *out = nullptr;
    const int avail = pt_device_count();
    if (avail == 0) return PT_ERR_NO_DEVICE;
    pt_sharded *S = new (std::nothrow) pt_sharded();
    if (!S) return PT_ERR_OUT_OF_MEMORY;
    try {
    const int R = opts->n_devices;
    S->slabs.resize(R);
    for (int r = 0; r < R; ++r) {
        S->slabs[r].device = opts->devices ? opts->devices[r] : r % avail;
        if (S->slabs[r].device < 0 || S->slabs[r].device >= avail) { delete S; return PT_ERR_INVALID_ARG; }
    }
    S->points = (const Rec80 *)points;
    S->n = n;
    S->coord_mode = opts->coord_mode;
    // x-cuts at point-count quantiles (of a sample of the cloud: the cuts need not be exact)
    {
        const size_t step = n > 4000000 ? n / 4000000 : 1;
        std::vector<double> xs;
        xs.reserve(n / step + 1);
        for (size_t i = 0; i < n; i += step) xs.push_back(S->points[i].ver[0]);
        S->cuts.assign(R + 1, 0.0);
        S->cuts[0] = -INFINITY;
        S->cuts[R] = INFINITY;
        for (int r = 1; r < R; ++r) {
            const size_t pos = xs.empty() ? 0 : std::min(xs.size() - 1, xs.size() * (size_t)r / R);
            if (!xs.empty()) std::nth_element(xs.begin(), xs.begin() + pos, xs.end());
            S->cuts[r] = xs.empty() ? 0.0 : xs[pos];
        }
        std::sort(S->cuts.begin() + 1, S->cuts.end() - 1);
    }
    // ghost-zone width: given, or a multiple of the expected k-th-neighbour distance estimated
    // from the bounding box (surface and volume estimate, the larger one); a call that finds it
    // too small widens it (sharded_query)
    S->halo = opts->halo;
    {
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (size_t i = 0; i < n; i += (n > 1000000 ? n / 1000000 : 1))
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], S->points[i].ver[a]); hi[a] = std::max(hi[a], S->points[i].ver[a]); }
        double e[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        std::sort(e, e + 3);
        const double kk = opts->k_hint > 0 ? opts->k_hint : 20;
        const double nn = n ? (double)n : 1.0;
        const double r_surface = std::sqrt(kk * std::max(e[1] * e[2], 0.0) / (M_PI * nn));
        const double r_volume = std::cbrt(3.0 * kk * std::max(e[0] * e[1] * e[2], 0.0) / (4.0 * M_PI * nn));
        S->halo_auto = 4.0 * std::max(r_surface, r_volume);
        if (!(S->halo_auto > 0.0) || !std::isfinite(S->halo_auto)) S->halo_auto = 1e-3;
        if (!(S->halo > 0.0)) S->halo = S->halo_auto;
    }
    const int rc = build_slabs(S);
    if (rc != PT_OK) { delete S; return rc; }
    *out = S;
    return PT_OK;
    } catch (...) {                 // host allocation failure: nothing is thrown across the ABI
        free_slabs(S);
        delete S;
        return PT_ERR_OUT_OF_MEMORY;
    }
}

int pt_sharded_free(pt_sharded *S)
{
    if (!S) return PT_OK;
    free_slabs(S);
    delete S;
    return PT_OK;
}

int pt_sharded_knn(pt_sharded *S, const void *queries, size_t m, int k, double radius, int32_t *idx_out,
                   double *d2_out)
{
    if (!idx_out && m) return PT_ERR_INVALID_ARG;
    try {
        return sharded_query(S, queries, m, k, radius, idx_out, d2_out, nullptr, nullptr);
    } catch (...) { return PT_ERR_OUT_OF_MEMORY; }
}

int pt_sharded_transfer(pt_sharded *S, const void *queries, size_t m, int k, double radius, int32_t *idx_out,
                        double *d2_out, uint8_t *rgba_out, float *normal_out)
{
    if ((!rgba_out && !normal_out) && m) return PT_ERR_INVALID_ARG;
    try {
        return sharded_query(S, queries, m, k, radius, idx_out, d2_out, rgba_out, normal_out);
    } catch (...) { return PT_ERR_OUT_OF_MEMORY; }
}

int pt_sharded_get_info(const pt_sharded *S, pt_sharded_info *info)
{
    if (!S || !info) return PT_ERR_INVALID_ARG;
    memset(info, 0, sizeof *info);
    info->n_slabs = (int)S->slabs.size();
    info->halo = S->halo;
    info->rebuilds = S->rebuilds;
    info->n_points = S->n;
    for (size_t r = 0; r < S->slabs.size() && r < 64; ++r) {
        info->slab_points[r] = S->slabs[r].n_own;
        info->slab_ghosts[r] = S->slabs[r].n_ghost;
        info->slab_device[r] = S->slabs[r].device;
    }
    return PT_OK;
}

}  // extern "C"
