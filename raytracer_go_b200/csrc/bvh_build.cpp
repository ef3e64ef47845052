// bvh_build.cpp — host build of the device BVH.  Replaces NewBVH (bvh.go:142-185): the reference
// splits on a random axis at the median of a descending sort; its topology differs on every run
// and is not part of the result, so this build is free to be a binned-SAH tree.  What must hold
// is that traversal returns World.Hit's answer (hittables.go:55-72), which needs box culling that
// can never reject a primitive the reference's float32 Hit (hittables.go:96-116, 167-190) accepts.
//
// Box padding.  With u = 2^-24, the float32 discriminant hittables.go:97-102 carries an absolute
// error of at most ~20 u |d|^2 |o-c|^2, so the reference can accept a root for a ray that
// geometrically passes up to  delta = 20 u |o-c|^2 / (2 r)  outside the sphere; the accepted
// point then lies within r + delta of the centre.  Each sphere's box is therefore grown by
//   pad = K u D^2 / (2 r) + 8 u (D + |m| + |c| + r),   K = 24,   D = r + origin_radius,
// i.e. for every ray that starts within `origin_radius` of the surface it hits (the second term
// covers the fused slab test's own rounding; m = median centre of the scene).  A quad's accepted point
// is within a few ulp of its plane and of its edges: its box (the reference's padded box,
// hittables.go:161 / bvh.go:63-82) is grown by 32 u (D + |Q| + |u| + |v|).  When the radius is derived
// (rt_scene_desc.ray_origin_radius = 0), rt_render / rt_trace enlarge it (refit, no rebuild) so that
// it covers the camera or the ray batch; an explicit radius is a fixed envelope.
#include "bvh_build.h"
#include "rt_shade.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <system_error>
#include <thread>

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; k++) lo[k] = std::numeric_limits<float>::infinity(), hi[k] = -lo[k];
    }
    void grow(const Box &b) {
        for (int k = 0; k < 3; k++) lo[k] = std::min(lo[k], b.lo[k]), hi[k] = std::max(hi[k], b.hi[k]);
    }
    double half_area() const {
        double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0;
        return dx * dy + dy * dz + dz * dx;
    }
};

const double U = 1.0 / 16777216.0; // 2^-24
const double K_DISC = 24.0;
const float NEG_INF = -std::numeric_limits<float>::infinity(), POS_INF = std::numeric_limits<float>::infinity();

// global primitive index g: [0, n_spheres) spheres, then quads
inline bool is_quad(const ScenePrims &p, uint32_t g) { return g >= p.spheres.size(); }

void prim_center(const ScenePrims &p, uint32_t g, double c[3], double *extent) {
    if (!is_quad(p, g)) {
        const rt_sphere &s = p.spheres[g];
        c[0] = s.cx, c[1] = s.cy, c[2] = s.cz;
        *extent = std::fabs((double)s.r);
    } else {
        const rt_quad &q = p.quads[g - p.spheres.size()];
        double e = 0;
        for (int k = 0; k < 3; k++) {
            c[k] = q.q[k] + 0.5 * ((double)q.u[k] + q.v[k]);
            e += 0.25 * ((double)q.u[k] + q.v[k]) * ((double)q.u[k] + q.v[k]);
        }
        double e2 = 0;
        for (int k = 0; k < 3; k++) e2 += 0.25 * ((double)q.u[k] - q.v[k]) * ((double)q.u[k] - q.v[k]);
        *extent = std::sqrt(std::max(e, e2));
    }
}

// Padded box of one primitive for rays that start within origin_radius of its surface.
Box padded_box(const ScenePrims &p, uint32_t g, const double m[3], double origin_radius, float *pad_out) {
    double c[3], ext;
    prim_center(p, g, c, &ext);
    const double D = origin_radius + ext; // bound on |o - c| (ext = radius, or a quad's half diagonal)
    const double cmax = std::max(std::fabs(c[0]), std::max(std::fabs(c[1]), std::fabs(c[2])));
    const double mmax = std::max(std::fabs(m[0]), std::max(std::fabs(m[1]), std::fabs(m[2])));
    Box b;
    double pad;
    if (!is_quad(p, g)) {
        const rt_sphere &s = p.spheres[g];
        const double r = std::fabs((double)s.r);
        pad = K_DISC * U * D * D / (2.0 * std::max(r, 1e-30)) + 8.0 * U * (D + mmax + cmax + r);
        pad = std::min(pad, D + r); // a box as large as the whole origin region is always enough
        const float cc[3] = {s.cx, s.cy, s.cz};
        for (int k = 0; k < 3; k++) {
            b.lo[k] = std::nextafter((float)((double)cc[k] - r - pad), NEG_INF);
            b.hi[k] = std::nextafter((float)((double)cc[k] + r + pad), POS_INF);
        }
    } else {
        const rt_quad &q = p.quads[g - p.spheres.size()];
        pad = 32.0 * U * (D + mmax + cmax + 2 * ext) + 1e-4; // >= the reference's own 1e-4 thin-box padding
        for (int k = 0; k < 3; k++) {
            const double a = q.q[k], e = (double)q.q[k] + q.u[k] + q.v[k];
            const double lo = std::min(std::min(a, e), std::min(a + q.u[k], a + q.v[k]));
            const double hi = std::max(std::max(a, e), std::max(a + q.u[k], a + q.v[k]));
            b.lo[k] = std::nextafter((float)(lo - pad), NEG_INF);
            b.hi[k] = std::nextafter((float)(hi + pad), POS_INF);
        }
    }
    *pad_out = (float)pad;
    return b;
}

// Device record of a quad: the fields NewQuad derives (hittables.go:149-165), float32, unfused.
void write_quad(const rt_quad &q, uint32_t id, F4 *out) {
    const V3 Q = v3(q.q[0], q.q[1], q.q[2]), u = v3(q.u[0], q.u[1], q.u[2]), v = v3(q.v[0], q.v[1], q.v[2]);
    const V3 n = v3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x); // Cross(u, v)
    const V3 normal = unit(n);
    const float D = dot(normal, Q);
    const V3 w = n * (1 / dot(n, n));
    out[0].x = Q.x, out[0].y = Q.y, out[0].z = Q.z, out[0].w = D;
    out[1].x = u.x, out[1].y = u.y, out[1].z = u.z, out[1].w = as_float(q.material);
    out[2].x = v.x, out[2].y = v.y, out[2].z = v.z, out[2].w = as_float(id);
    out[3].x = w.x, out[3].y = w.y, out[3].z = w.z, out[3].w = 0;
    out[4].x = normal.x, out[4].y = normal.y, out[4].z = normal.z, out[4].w = 0;
}

// The build runs in two parallel phases over one shared thread budget (RT_B200_BVH_THREADS, default
// min(16, hardware threads)):
//   split()  recursive binned-SAH partitioning of the primitive records into a temporary tree; subtrees above
//            PAR_SUBTREE primitives run as their own threads, and the bounds / binning passes of a range
//            above PAR_RANGE are chunked over idle threads (min/max and counts merge exactly);
//   emit()   writes the flat arrays.  Every temporary node knows how many device nodes and slots its
//            subtree takes, so a subtree's place in the depth-first layout is known before it is written
//            and subtrees are emitted concurrently.
// Both phases are deterministic: the arrays are identical for any thread count (tested), and identical
// to what a single depth-first recursion produces.
struct TNode {
    Box box;
    uint32_t left, right; // pool indices; left == TN_LEAF: a leaf over items[b, e)
    uint32_t b, e;
    uint32_t n_nodes, n_sph, n_quad; // device nodes / sphere slots / quad slots of the subtree
    uint32_t height;                 // chain of inner nodes below and including this one
};
const uint32_t TN_LEAF = 0xFFFFFFFFu;

struct ThreadBudget {
    std::atomic<int> idle{0};
    int take(int want) { // up to `want` helper threads
        int got = 0;
        while (got < want) {
            int cur = idle.load(std::memory_order_relaxed);
            if (cur <= 0) break;
            if (idle.compare_exchange_weak(cur, cur - 1)) got++;
        }
        return got;
    }
    void give(int n) { idle.fetch_add(n); }
};

// Runs fn on a new thread; if the thread cannot be created (process / cgroup limits) runs it on the
// caller instead, so a build never throws through the C ABI.
template <class F>
void spawn_or_run(std::vector<std::thread> &th, F fn) {
    try {
        th.emplace_back(fn);
    } catch (const std::system_error &) {
        fn();
    }
}

struct Builder {
    const ScenePrims *prims;
    // One record per primitive, partitioned in place: every pass over a range is a linear scan
    // (partitioning an index array instead made the top levels of a 1 M build cache-miss bound).
    struct Prim {
        Box box;    // padded box
        float c[3]; // centre
        uint32_t g; // global primitive index
    };
    RawVec<Prim> items;
    FlatBvh *out;
    int max_leaf;
    std::unique_ptr<TNode[]> pool; // uninitialised; nodes are handed out in per-thread blocks
    size_t pool_size = 0;
    std::atomic<uint32_t> pool_next{0};
    static const uint32_t POOL_BLOCK = 256;
    uint64_t build_id = next_build_id(); // unique per Builder, never 0
    static uint64_t next_build_id() {
        static std::atomic<uint64_t> id{0};
        return ++id;
    }
    ThreadBudget budget;
    int threads = 1;

    static const int NBINS = 64; // upper bound; `nbins` bins are used
    int nbins_max = 64; // measured on C2: 16 -> 64 bins = 3 % fewer box tests per ray, +2.5 % Msamples/s
    double c_trav = 1.2; // cost of visiting a node pair, in primitive tests
    static const size_t PAR_SUBTREE = 16384; // smallest subtree worth its own thread
    static const size_t PAR_RANGE = 131072;  // smallest range whose passes are chunked over threads
    static const uint32_t DEPTH_MEDIAN = 30; // inner-node depth from which ranges are split by count

    struct RangeInfo {
        Box bounds, cb;
        bool any_quad = false, any_sphere = false;
        void reset() { bounds.reset(), cb.reset(), any_quad = any_sphere = false; }
        void merge(const RangeInfo &o) {
            bounds.grow(o.bounds), cb.grow(o.cb);
            any_quad |= o.any_quad, any_sphere |= o.any_sphere;
        }
    };
    struct Bins {
        Box box[3][NBINS];
        size_t n[3][NBINS];
        void reset(int nbins) {
            for (int a = 0; a < 3; a++)
                for (int k = 0; k < nbins; k++) box[a][k].reset(), n[a][k] = 0;
        }
        void merge(const Bins &o, int nbins) {
            for (int a = 0; a < 3; a++)
                for (int k = 0; k < nbins; k++) box[a][k].grow(o.box[a][k]), n[a][k] += o.n[a][k];
        }
    };

    void range_info(size_t b, size_t e, RangeInfo *r) const {
        r->reset();
        for (size_t i = b; i < e; i++) {
            const Prim &it = items[i];
            r->bounds.grow(it.box);
            (is_quad(*prims, it.g) ? r->any_quad : r->any_sphere) = true;
            for (int k = 0; k < 3; k++) r->cb.lo[k] = std::min(r->cb.lo[k], it.c[k]), r->cb.hi[k] = std::max(r->cb.hi[k], it.c[k]);
        }
    }
    // centroid bins of all three axes in one pass over the range (an axis of zero extent is skipped)
    void bin_range(size_t b, size_t e, const Box &cb, int nbins, Bins *bins) const {
        bins->reset(nbins);
        float lo[3], scale[3];
        bool use[3];
        for (int a = 0; a < 3; a++) {
            const float ext = cb.hi[a] - cb.lo[a];
            lo[a] = cb.lo[a], use[a] = ext > 0, scale[a] = use[a] ? (float)nbins / ext : 0.0f;
        }
        for (size_t i = b; i < e; i++) {
            const Prim &it = items[i];
            const Box &bx = it.box;
            for (int a = 0; a < 3; a++) {
                if (!use[a]) continue;
                const int k = std::min(nbins - 1, std::max(0, (int)((it.c[a] - lo[a]) * scale[a])));
                bins->box[a][k].grow(bx);
                bins->n[a][k]++;
            }
        }
    }
    // fn(chunk_begin, chunk_end, chunk_index) over [b, e) on 1 + helpers threads
    template <class F>
    void chunked(size_t b, size_t e, int helpers, F fn) {
        const int parts = helpers + 1;
        const size_t step = (e - b + parts - 1) / parts;
        std::vector<std::thread> th;
        for (int c = 1; c < parts; c++) {
            const size_t cb_ = std::min(e, b + c * step), ce = std::min(e, cb_ + step);
            spawn_or_run(th, [=, &fn] { fn(cb_, ce, c); });
        }
        fn(b, std::min(e, b + step), 0);
        for (auto &t : th) t.join();
    }

    uint32_t new_node() {
        static thread_local uint32_t next = 0, end = 0;
        static thread_local uint64_t owner = 0; // build_id of the build the block belongs to
        if (owner != build_id || next == end) {
            next = pool_next.fetch_add(POOL_BLOCK), end = next + POOL_BLOCK, owner = build_id;
            if (end > pool_size) abort(); // sized for the worst case in build_flat_bvh
        }
        return next++;
    }

    // phase 1: the subtree over items[b, e) as a temporary tree; returns its pool index
    uint32_t split(size_t b, size_t e, uint32_t depth = 0) {
        const size_t n = e - b;
        const uint32_t me = new_node();
        // helper threads in proportion to the range's share of the build, so that sibling ranges get equal help
        int helpers = n >= PAR_RANGE ? budget.take((int)((double)threads * (double)n / (double)items.size()) - 1) : 0;
        RangeInfo info;
        if (helpers) {
            std::vector<RangeInfo> part(helpers + 1);
            chunked(b, e, helpers, [&](size_t cb_, size_t ce, int c) { range_info(cb_, ce, &part[c]); });
            info.reset();
            for (auto &p : part) info.merge(p);
        } else {
            range_info(b, e, &info);
        }
        const Box &bounds = info.bounds, &cb = info.cb;
        const bool any_quad = info.any_quad, any_sphere = info.any_sphere;
        const int nbins = (int)std::min<size_t>((size_t)nbins_max, std::max<size_t>(4, n)); // no more bins than primitives
        // SAH over centroid bins per axis; cost unit = one primitive test
        double best_cost = std::numeric_limits<double>::infinity();
        int best_axis = -1, best_bin = -1;
        if (n > 1) {
            std::unique_ptr<Bins> bins(new Bins);
            if (helpers) {
                std::vector<std::unique_ptr<Bins>> part(helpers + 1);
                for (auto &p : part) p.reset(new Bins);
                chunked(b, e, helpers, [&](size_t cb_, size_t ce, int c) { bin_range(cb_, ce, cb, nbins, part[c].get()); });
                bins->reset(nbins);
                for (auto &p : part) bins->merge(*p, nbins);
            } else {
                bin_range(b, e, cb, nbins, bins.get());
            }
            for (int axis = 0; axis < 3; axis++) {
                if (!(cb.hi[axis] - cb.lo[axis] > 0)) continue;
                const Box *bin_box = bins->box[axis];
                const size_t *bin_n = bins->n[axis];
                double right_area[NBINS];
                size_t right_n[NBINS];
                Box acc;
                acc.reset();
                size_t cnt = 0;
                for (int k = nbins - 1; k > 0; k--) {
                    acc.grow(bin_box[k]);
                    cnt += bin_n[k];
                    right_area[k] = acc.half_area(), right_n[k] = cnt;
                }
                acc.reset();
                cnt = 0;
                for (int k = 0; k < nbins - 1; k++) {
                    acc.grow(bin_box[k]);
                    cnt += bin_n[k];
                    if (cnt == 0 || right_n[k + 1] == 0) continue;
                    double cost = acc.half_area() * (double)cnt + right_area[k + 1] * (double)right_n[k + 1];
                    if (cost < best_cost) best_cost = cost, best_axis = axis, best_bin = k;
                }
            }
        }
        if (helpers) budget.give(helpers);
        const double parent_area = std::max(bounds.half_area(), 1e-30);
        const double split_cost = c_trav + best_cost / parent_area;
        TNode &t = pool[me];
        t.box = bounds, t.b = (uint32_t)b, t.e = (uint32_t)e;
        // a leaf holds primitives of one kind; quads are one per leaf (80-byte records)
        const bool can_leaf = !(any_quad && any_sphere) && n <= (size_t)(any_quad ? 1 : max_leaf);
        if (can_leaf && (best_axis < 0 || (double)n <= split_cost)) {
            t.left = t.right = TN_LEAF;
            t.n_nodes = 0, t.height = 0;
            t.n_sph = any_quad ? 0 : (uint32_t)n, t.n_quad = any_quad ? (uint32_t)n : 0;
            if (!any_quad) // ascending object ID inside a leaf
                std::sort(items.begin() + b, items.begin() + e,
                          [&](const Prim &x, const Prim &y) { return prims->sphere_ids[x.g] < prims->sphere_ids[y.g]; });
            return me;
        }

        size_t mid;
        if (depth >= DEPTH_MEDIAN) {
            // SAH can peel one primitive per level off a scene whose extents grow geometrically; from this
            // depth on the ranges are halved by count along their widest axis instead, so the chain of inner
            // nodes stays below DEPTH_MEDIAN + 32 <= the kernels' traversal stack (RT_LOCAL_STACK)
            int ax = 0;
            for (int k = 1; k < 3; k++)
                if (cb.hi[k] - cb.lo[k] > cb.hi[ax] - cb.lo[ax]) ax = k;
            mid = b + n / 2;
            std::nth_element(items.begin() + b, items.begin() + mid, items.begin() + e, [ax](const Prim &x, const Prim &y) {
                return x.c[ax] < y.c[ax] || (x.c[ax] == y.c[ax] && x.g < y.g);
            });
        } else if (best_axis >= 0) {
            float lo = cb.lo[best_axis], ext = cb.hi[best_axis] - cb.lo[best_axis];
            float scale = (float)nbins / ext;
            auto it = std::partition(items.begin() + b, items.begin() + e, [&](const Prim &q) {
                int k = std::min(nbins - 1, std::max(0, (int)((q.c[best_axis] - lo) * scale)));
                return k <= best_bin;
            });
            mid = (size_t)(it - items.begin());
        } else if (any_quad && any_sphere) {
            auto it = std::partition(items.begin() + b, items.begin() + e, [&](const Prim &q) { return !is_quad(*prims, q.g); });
            mid = (size_t)(it - items.begin());
        } else {
            mid = b + n / 2; // coincident centres: split by count
        }
        if (mid == b || mid == e) mid = b + n / 2;

        uint32_t l, r;
        if (std::min(mid - b, e - mid) >= PAR_SUBTREE && budget.take(1)) {
            std::vector<std::thread> th;
            spawn_or_run(th, [&] { l = split(b, mid, depth + 1); });
            r = split(mid, e, depth + 1);
            for (auto &t : th) t.join();
            budget.give(1);
        } else {
            l = split(b, mid, depth + 1);
            r = split(mid, e, depth + 1);
        }
        TNode &tt = pool[me]; // (the pool never reallocates)
        const TNode &tl = pool[l], &tr = pool[r];
        tt.left = l, tt.right = r;
        tt.n_nodes = 2 + tl.n_nodes + tr.n_nodes;
        tt.n_sph = tl.n_sph + tr.n_sph, tt.n_quad = tl.n_quad + tr.n_quad;
        tt.height = 1 + std::max(tl.height, tr.height);
        return me;
    }

    // phase 2: write the subtree of pool node `ti` — its device nodes start at node index `node_off`
    // (parents precede children, siblings adjacent: depth-first order), its slots at sph_off / quad_off
    uint32_t emit(uint32_t ti, uint32_t node_off, uint32_t sph_off, uint32_t quad_off) {
        const TNode &t = pool[ti];
        if (t.left == TN_LEAF) return emit_leaf(t, sph_off, quad_off);
        const TNode &tl = pool[t.left], &tr = pool[t.right];
        const uint32_t pair = node_off;
        uint32_t lref, rref;
        const uint32_t r_node = node_off + 2 + tl.n_nodes, r_sph = sph_off + tl.n_sph, r_quad = quad_off + tl.n_quad;
        if (std::min(tl.n_sph + tl.n_quad, tr.n_sph + tr.n_quad) >= PAR_SUBTREE && budget.take(1)) {
            std::vector<std::thread> th;
            spawn_or_run(th, [&] { lref = emit(t.left, node_off + 2, sph_off, quad_off); });
            rref = emit(t.right, r_node, r_sph, r_quad);
            for (auto &t : th) t.join();
            budget.give(1);
        } else {
            lref = emit(t.left, node_off + 2, sph_off, quad_off);
            rref = emit(t.right, r_node, r_sph, r_quad);
        }
        write_node(pair, tl.box, lref);
        write_node(pair + 1, tr.box, rref);
        return pair;
    }

    void write_node(uint32_t i, const Box &bx, uint32_t ref) {
        F4 a, c;
        a.x = bx.lo[0], a.y = bx.lo[1], a.z = bx.lo[2];
        c.x = bx.hi[0], c.y = bx.hi[1], c.z = bx.hi[2], c.w = 0;
        memcpy(&a.w, &ref, 4);
        out->nodes[2 * (size_t)i] = a, out->nodes[2 * (size_t)i + 1] = c;
    }

    uint32_t emit_leaf(const TNode &t, uint32_t sph_off, uint32_t quad_off) {
        const ScenePrims &p = *prims;
        const size_t b = t.b, e = t.e;
        if (is_quad(p, items[b].g)) {
            const uint32_t first = quad_off;
            for (size_t i = b; i < e; i++) {
                const uint32_t qi = items[i].g - (uint32_t)p.spheres.size();
                const size_t slot = (size_t)quad_off + (i - b);
                write_quad(p.quads[qi], p.quad_ids[qi], &out->quad[(size_t)RT_QUAD_F4 * slot]);
                out->quad_prim[slot] = qi;
            }
            return RT_LEAF | RT_LEAF_QUAD | (first << 3) | (uint32_t)(e - b - 1);
        }
        const uint32_t first = sph_off;
        for (size_t i = b; i < e; i++) {
            const uint32_t g = items[i].g;
            const rt_sphere &s = p.spheres[g];
            const size_t slot = (size_t)sph_off + (i - b);
            F4 f;
            f.x = s.cx, f.y = s.cy, f.z = s.cz, f.w = s.r;
            out->sph[slot] = f;
            I2 m;
            m.x = (int32_t)p.sphere_ids[g], m.y = (int32_t)s.material;
            out->meta[slot] = m;
            out->sph_prim[slot] = g;
        }
        return RT_LEAF | (first << 3) | (uint32_t)(e - b - 1);
    }
};

} // namespace

bool load_scene_prims(const rt_scene_desc *d, ScenePrims *out) {
    out->spheres.assign(d->spheres, d->spheres + d->n_spheres);
    out->quads.assign(d->quads, d->quads + d->n_quads);
    const size_t ns = out->spheres.size(), nq = out->quads.size(), n = ns + nq;
    out->sphere_ids.resize(ns), out->quad_ids.resize(nq);
    for (size_t i = 0; i < ns; i++) out->sphere_ids[i] = d->sphere_ids ? d->sphere_ids[i] : (uint32_t)i;
    for (size_t i = 0; i < nq; i++) out->quad_ids[i] = d->quad_ids ? d->quad_ids[i] : (uint32_t)(ns + i);
    if (d->sphere_ids || d->quad_ids) { // must be a permutation of 0..n-1
        std::vector<bool> seen(n, false);
        for (size_t i = 0; i < n; i++) {
            const uint32_t id = i < ns ? out->sphere_ids[i] : out->quad_ids[i - ns];
            if (id >= n || seen[id]) return false;
            seen[id] = true;
        }
    }
    return true;
}

void compute_scene_center(const ScenePrims &prims, double m[3], double *extent90, double *surface_extent) {
    m[0] = m[1] = m[2] = 0, *extent90 = 0;
    if (surface_extent) *surface_extent = 0;
    const size_t n = prims.size();
    if (n == 0) return;
    std::vector<float> v(n);
    std::vector<double> c(3 * n), ext(n);
    for (size_t g = 0; g < n; g++) prim_center(prims, (uint32_t)g, &c[3 * g], &ext[g]);
    for (int k = 0; k < 3; k++) {
        for (size_t g = 0; g < n; g++) v[g] = (float)c[3 * g + k];
        std::nth_element(v.begin(), v.begin() + n / 2, v.end());
        m[k] = v[n / 2];
    }
    for (size_t g = 0; g < n; g++) {
        double dx = c[3 * g] - m[0], dy = c[3 * g + 1] - m[1], dz = c[3 * g + 2] - m[2];
        const double dc = std::sqrt(dx * dx + dy * dy + dz * dz);
        v[g] = (float)(dc + ext[g]);
        // how far from m the nearest point of the farthest primitive can be
        if (surface_extent) *surface_extent = std::max(*surface_extent, dc - ext[g]);
    }
    size_t k90 = (size_t)((n - 1) * 0.9);
    std::nth_element(v.begin(), v.begin() + k90, v.end());
    *extent90 = v[k90];
}

static int build_threads() {
    int t = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char *e = getenv("RT_B200_BVH_THREADS")) t = std::max(1, std::min(64, atoi(e)));
    return t;
}

// fn(begin, end) over [0, n) on up to `threads` threads (small ranges stay on the caller)
template <class F>
static void parallel_ranges(size_t n, int threads, F fn) {
    const int parts = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, n / 65536));
    if (parts <= 1) {
        fn((size_t)0, n);
        return;
    }
    const size_t step = (n + parts - 1) / parts;
    std::vector<std::thread> th;
    for (int c = 1; c < parts; c++) spawn_or_run(th, [=, &fn] { fn(std::min(n, c * step), std::min(n, (c + 1) * step)); });
    fn((size_t)0, std::min(n, step));
    for (auto &t : th) t.join();
}

// (min, max) boxes -> the (centre, half-extent) form the traversal kernels read.  The centre is rounded
// to float; the half-extent is taken from that rounded centre, rounded up, and widened by a few ulp of
// the magnitudes involved for the one extra rounding the centre form has in the slab test
// (m = (c - o) * inv is rounded before -+ h*|inv| is added).  Culling only has to be conservative.
static void make_device_nodes(FlatBvh *bvh) {
    const double U = 5.960464477539063e-08; // 2^-24
    bvh->dev_nodes.resize(bvh->nodes.size());
    parallel_ranges(bvh->nodes.size() / 2, build_threads(), [&](size_t n0, size_t n1) {
    for (size_t i = 2 * n0; i < 2 * n1; i += 2) {
        const F4 &lo = bvh->nodes[i], &hi = bvh->nodes[i + 1];
        const float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
        float c[3], e[3];
        for (int k = 0; k < 3; k++) {
            c[k] = (float)(0.5 * ((double)l[k] + (double)h[k]));
            double half = std::max((double)h[k] - (double)c[k], (double)c[k] - (double)l[k]);
            half += 4.0 * U * (std::fabs((double)c[k]) + half);
            e[k] = std::nextafter((float)half, POS_INF);
            if (!((double)c[k] - (double)e[k] <= (double)l[k] && (double)c[k] + (double)e[k] >= (double)h[k])) e[k] = POS_INF; // overflow
        }
        F4 a = {c[0], c[1], c[2], lo.w}, b = {e[0], e[1], e[2], 0.0f};
        bvh->dev_nodes[i] = a, bvh->dev_nodes[i + 1] = b;
    }
    });
    // walk pairs of the leaf-start chains: copies of two real nodes side by side (see bvh_build.h)
    const size_t n_real = bvh->nodes.size(), n_walk = bvh->walk_src.size() / 2;
    bvh->dev_nodes.resize(n_real + 4 * n_walk);
    for (size_t k = 0; k < n_walk; k++) {
        F4 *dst = &bvh->dev_nodes[n_real + 4 * k];
        const uint32_t s0 = bvh->walk_src[2 * k], s1 = bvh->walk_src[2 * k + 1];
        dst[0] = bvh->dev_nodes[2 * (size_t)s0], dst[1] = bvh->dev_nodes[2 * (size_t)s0 + 1];
        if (s1 != RT_REF_NONE) {
            dst[2] = bvh->dev_nodes[2 * (size_t)s1], dst[3] = bvh->dev_nodes[2 * (size_t)s1 + 1];
        } else {
            // no second sibling: a box that cannot be hit (far plane in front of the near plane on every axis);
            // should the slab test ever degenerate (zero reciprocal), its ref repeats the first box's, which is harmless
            dst[2] = dst[0], dst[3] = dst[1];
            dst[2].x = dst[2].y = dst[2].z = 0.0f;
            dst[3].x = dst[3].y = dst[3].z = -1e30f;
        }
    }
}

// Leaf-start chains (bvh_build.h).  Serial; called for scenes small enough to be staged in shared memory.
static void build_leaf_start(FlatBvh *bvh) {
    const size_t n_nodes = bvh->nodes.size() / 2;
    bvh->walk_src.clear(), bvh->chains.clear();
    bvh->sph_chain.assign(bvh->sph.size(), RT_REF_NONE);
    bvh->quad_chain.assign(bvh->quad_prim.size(), RT_REF_NONE);
    if (n_nodes == 0) return;
    auto ref_of = [&](size_t i) {
        uint32_t r;
        memcpy(&r, &bvh->nodes[2 * i].w, 4);
        return r;
    };
    std::vector<uint32_t> parent(n_nodes / 2, RT_REF_NONE); // per pair: the node whose ref it is
    for (size_t i = 0; i < n_nodes; i++)
        if (!(ref_of(i) & RT_LEAF)) parent[ref_of(i) / 2] = (uint32_t)i;
    std::vector<uint32_t> pair_of(n_nodes, RT_REF_NONE); // walk pair whose first box is the sibling of node i
    std::vector<uint32_t> chain;
    for (size_t leaf = 0; leaf < n_nodes; leaf++) {
        const uint32_t lref = ref_of(leaf);
        if (!(lref & RT_LEAF)) continue;
        chain.clear();
        for (uint32_t n = (uint32_t)leaf;;) {
            const uint32_t up = parent[n / 2]; // parent of n (RT_REF_NONE: n is a child of the root)
            if (pair_of[n] == RT_REF_NONE) {
                pair_of[n] = (uint32_t)(n_nodes + bvh->walk_src.size()); // 2 device nodes per walk pair, appended after the real ones
                bvh->walk_src.push_back(n ^ 1u);
                bvh->walk_src.push_back(up == RT_REF_NONE ? RT_REF_NONE : (up ^ 1u));
            }
            chain.push_back(pair_of[n]);
            if (up == RT_REF_NONE) break;
            n = parent[up / 2];
            if (n == RT_REF_NONE) break;
        }
        const uint32_t off = (uint32_t)bvh->chains.size();
        bvh->chains.push_back((uint32_t)chain.size());
        for (size_t k = chain.size(); k-- > 0;) bvh->chains.push_back(chain[k]); // root side first: the deepest pair is popped first
        bvh->chains.push_back(lref);
        const uint32_t first = (lref & RT_LEAF_SLOT_MASK) >> 3, count = (lref & 7u) + 1;
        for (uint32_t q = first; q < first + count; q++) ((lref & RT_LEAF_QUAD) ? bvh->quad_chain : bvh->sph_chain)[q] = off;
    }
}

void build_flat_bvh(const ScenePrims &prims, float origin_radius, int max_leaf, FlatBvh *out, const double *center,
                    bool leaf_start) {
    *out = FlatBvh();
    const size_t n = prims.size();
    if (n == 0) return;
    max_leaf = std::max(1, std::min(max_leaf, RT_MAX_LEAF));
    const int threads = build_threads();
    Builder b;
    b.prims = &prims, b.out = out, b.max_leaf = max_leaf;
    if (const char *e = getenv("RT_B200_BVH_BINS")) b.nbins_max = std::max(2, std::min(64, atoi(e)));
    if (const char *e = getenv("RT_B200_BVH_CTRAV")) b.c_trav = atof(e);
    b.items.resize(n);
    double m[3], ext;
    const bool timing = getenv("RT_B200_BVH_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[bvh] %-14s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    if (center)
        m[0] = center[0], m[1] = center[1], m[2] = center[2];
    else
        compute_scene_center(prims, m, &ext);
    lap("scene centre");
    std::vector<float> part_min(threads + 1, POS_INF), part_max(threads + 1, 0.0f);
    std::atomic<int> part_next{0};
    parallel_ranges(n, threads, [&](size_t g0, size_t g1) {
        float pmin = POS_INF, pmax = 0;
        for (size_t g = g0; g < g1; g++) {
            float pad;
            Builder::Prim &it = b.items[g];
            it.box = padded_box(prims, (uint32_t)g, m, origin_radius, &pad);
            pmin = std::min(pmin, pad), pmax = std::max(pmax, pad);
            double c[3], e;
            prim_center(prims, (uint32_t)g, c, &e);
            it.c[0] = (float)c[0], it.c[1] = (float)c[1], it.c[2] = (float)c[2];
            it.g = (uint32_t)g;
        }
        const int slot = part_next.fetch_add(1);
        part_min[slot] = pmin, part_max[slot] = pmax;
    });
    out->pad_min = *std::min_element(part_min.begin(), part_min.end());
    out->pad_max = *std::max_element(part_max.begin(), part_max.end());

    lap("padded boxes");
    // a binary tree over n primitives has at most 2n - 1 nodes; every thread may leave one block partly used
    b.pool_size = 2 * n + (2 * n / Builder::PAR_SUBTREE + 64) * Builder::POOL_BLOCK;
    b.pool.reset(new TNode[b.pool_size]);
    b.budget.idle = threads - 1, b.threads = threads;
    lap("pool");
    const uint32_t root = b.split(0, n);
    lap("split");
    const TNode &rt = b.pool[root];
    out->nodes.resize(2 * (size_t)rt.n_nodes);
    out->sph.resize(rt.n_sph), out->meta.resize(rt.n_sph), out->sph_prim.resize(rt.n_sph);
    out->quad.resize((size_t)RT_QUAD_F4 * rt.n_quad), out->quad_prim.resize(rt.n_quad);
    out->max_depth = rt.height;
    lap("resize");
    out->root_ref = b.emit(root, 0, 0, 0);
    lap("emit");
    if (leaf_start) build_leaf_start(out), lap("leaf start");
    make_device_nodes(out);
    lap("device nodes");
}

// Recompute every box for a larger origin_radius, topology unchanged.  Nodes are in pre-order
// (children after parents), so one reverse sweep rebuilds parents from children.
void refit_flat_bvh(const ScenePrims &prims, float origin_radius, FlatBvh *bvh) {
    if (prims.size() == 0 || bvh->root_ref == RT_REF_NONE) return;
    double m[3], ext;
    compute_scene_center(prims, m, &ext);
    const size_t n_nodes = bvh->nodes.size() / 2;
    const int threads = build_threads();
    auto store = [&](size_t ii, const Box &bx) {
        F4 &lo = bvh->nodes[2 * ii], &hi = bvh->nodes[2 * ii + 1];
        lo.x = bx.lo[0], lo.y = bx.lo[1], lo.z = bx.lo[2];
        hi.x = bx.hi[0], hi.y = bx.hi[1], hi.z = bx.hi[2];
    };
    // pass 1 (parallel): leaf boxes from their primitives' padded boxes
    std::vector<float> part_min(threads + 1, POS_INF), part_max(threads + 1, 0.0f);
    std::atomic<int> part_next{0};
    parallel_ranges(n_nodes, threads, [&](size_t n0, size_t n1) {
        float pmin = POS_INF, pmax = 0;
        for (size_t ii = n0; ii < n1; ii++) {
            uint32_t ref;
            memcpy(&ref, &bvh->nodes[2 * ii].w, 4);
            if (!(ref & RT_LEAF)) continue;
            Box bx;
            bx.reset();
            const bool quad = (ref & RT_LEAF_QUAD) != 0;
            uint32_t first = (ref & RT_LEAF_SLOT_MASK) >> 3, count = (ref & 7u) + 1;
            for (uint32_t s = first; s < first + count; s++) {
                float pad;
                const uint32_t g = quad ? (uint32_t)prims.spheres.size() + bvh->quad_prim[s] : bvh->sph_prim[s];
                bx.grow(padded_box(prims, g, m, origin_radius, &pad));
                pmin = std::min(pmin, pad), pmax = std::max(pmax, pad);
            }
            store(ii, bx);
        }
        const int slot = part_next.fetch_add(1);
        part_min[slot] = pmin, part_max[slot] = pmax;
    });
    // pass 2: inner boxes, children before parents (reverse pre-order)
    for (size_t ii = n_nodes; ii-- > 0;) {
        uint32_t ref;
        memcpy(&ref, &bvh->nodes[2 * ii].w, 4);
        if (ref & RT_LEAF) continue;
        Box bx;
        bx.reset();
        for (int c = 0; c < 2; c++) {
            const F4 &lo = bvh->nodes[2 * ((size_t)ref + c)], &hi = bvh->nodes[2 * ((size_t)ref + c) + 1];
            Box cbx;
            cbx.lo[0] = lo.x, cbx.lo[1] = lo.y, cbx.lo[2] = lo.z;
            cbx.hi[0] = hi.x, cbx.hi[1] = hi.y, cbx.hi[2] = hi.z;
            bx.grow(cbx);
        }
        store(ii, bx);
    }
    bvh->pad_min = *std::min_element(part_min.begin(), part_min.end());
    bvh->pad_max = *std::max_element(part_max.begin(), part_max.end());
    make_device_nodes(bvh);
}

void pack_materials(const rt_scene_desc *d, std::vector<F4> *out) {
    out->resize(2 * (size_t)d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const rt_material &m = d->materials[i];
        F4 m0 = {0, 0, 0, 0}, m1 = {0, 0, 0, 0};
        uint32_t code = RT_CODE(m.kind, 0, 0);
        if (m.kind == RT_MAT_METAL) {
            m0.x = m.albedo[0], m0.y = m.albedo[1], m0.z = m.albedo[2], m0.w = m.fuzz;
        } else if (m.kind == RT_MAT_DIELECTRIC) {
            m0.w = m.ior;
            m1.x = 1.0f / m.ior; // materials.go:94
        } else {
            const rt_texture &t = d->textures[m.texture];
            code = RT_CODE(m.kind, t.kind, t.kind == RT_TEX_IMAGE ? t.image : 0);
            if (t.kind == RT_TEX_CHECKER) {
                m0.x = t.a[0], m0.y = t.a[1], m0.z = t.a[2];
                m0.w = 1 / t.scale; // materials.go:128
                m1.x = t.b[0], m1.y = t.b[1], m1.z = t.b[2];
            } else if (t.kind == RT_TEX_NOISE) {
                code = RT_CODE(m.kind, t.kind, t.image); // index into the Perlin tables
                m0.w = t.scale;                            // NoiseTexture.scale, materials.go:282
            } else if (t.kind == RT_TEX_IMAGE) {
                m0.x = t.oob[0], m0.y = t.oob[1], m0.z = t.oob[2];
            } else {
                m0.x = t.a[0], m0.y = t.a[1], m0.z = t.a[2];
            }
        }
        memcpy(&m1.w, &code, 4);
        (*out)[2 * (size_t)i] = m0, (*out)[2 * (size_t)i + 1] = m1;
    }
}

