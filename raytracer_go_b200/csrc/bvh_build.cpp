// bvh_build.cpp — host build of the device BVH.  Replaces NewBVH (bvh.go:142-185): the reference
// splits on a random axis at the median of a descending sort; its topology differs on every run
// and is not part of the result, so this build is free to be a binned-SAH tree.  What must hold
// is that traversal returns World.Hit's answer (hittables.go:55-72), which needs box culling that
// can never reject a sphere the reference's float32 Sphere.Hit (hittables.go:96-116) accepts.
//
// Box padding.  With u = 2^-24, the float32 discriminant hittables.go:97-102 carries an absolute
// error of at most ~20 u |d|^2 |o-c|^2, so the reference can accept a root for a ray that
// geometrically passes up to  delta = 20 u |o-c|^2 / (2 r)  outside the sphere; the accepted
// point then lies within r + delta of the centre.  Each sphere's box is therefore grown by
//   pad = K u D^2 / (2 r) + 8 u (D + |m| + |c| + r),   K = 24,
// where D bounds |o - c| for every ray origin o within `origin_radius` of the scene's median
// centre m (the second term covers the fused slab test's own rounding).  rt_render / rt_trace
// enlarge origin_radius (refit, no rebuild) when a camera or a ray batch lies outside it.
#include "bvh_build.h"
#include "rt_shade.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

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

struct Builder {
    const rt_sphere *sph;
    std::vector<Box> boxes;      // padded per-sphere boxes
    std::vector<uint32_t> order; // permutation being partitioned
    FlatBvh *out;
    int max_leaf;

    static const int NBINS = 64; // upper bound; `nbins` bins are used
    int nbins = 16;
    double c_trav = 1.2; // cost of visiting a node pair, in sphere tests

    // returns the ref of the subtree over order[b, e), writes its box
    uint32_t build(size_t b, size_t e, uint32_t depth, Box *box_out) {
        Box bounds, cb;
        bounds.reset(), cb.reset();
        for (size_t i = b; i < e; i++) {
            const Box &bx = boxes[order[i]];
            bounds.grow(bx);
            const rt_sphere &s = sph[order[i]];
            float c[3] = {s.cx, s.cy, s.cz};
            for (int k = 0; k < 3; k++) cb.lo[k] = std::min(cb.lo[k], c[k]), cb.hi[k] = std::max(cb.hi[k], c[k]);
        }
        *box_out = bounds;
        const size_t n = e - b;
        // SAH over 16 centroid bins per axis; cost unit = one sphere test, a box pair costs 1.2
        double best_cost = std::numeric_limits<double>::infinity();
        int best_axis = -1, best_bin = -1;
        if (n > 1) {
            for (int axis = 0; axis < 3; axis++) {
                float lo = cb.lo[axis], ext = cb.hi[axis] - cb.lo[axis];
                if (!(ext > 0)) continue;
                Box bin_box[NBINS];
                size_t bin_n[NBINS] = {0};
                for (int k = 0; k < nbins; k++) bin_box[k].reset();
                float scale = (float)nbins / ext;
                for (size_t i = b; i < e; i++) {
                    const rt_sphere &s = sph[order[i]];
                    float c = axis == 0 ? s.cx : axis == 1 ? s.cy : s.cz;
                    int k = std::min(nbins - 1, std::max(0, (int)((c - lo) * scale)));
                    bin_box[k].grow(boxes[order[i]]);
                    bin_n[k]++;
                }
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
        const double parent_area = std::max(bounds.half_area(), 1e-30);
        const double split_cost = c_trav + best_cost / parent_area;
        const bool can_leaf = n <= (size_t)max_leaf;
        if (can_leaf && (best_axis < 0 || (double)n <= split_cost)) return make_leaf(b, e);

        size_t mid;
        if (best_axis >= 0) {
            float lo = cb.lo[best_axis], ext = cb.hi[best_axis] - cb.lo[best_axis];
            float scale = (float)nbins / ext;
            auto it = std::partition(order.begin() + b, order.begin() + e, [&](uint32_t p) {
                const rt_sphere &s = sph[p];
                float c = best_axis == 0 ? s.cx : best_axis == 1 ? s.cy : s.cz;
                int k = std::min(nbins - 1, std::max(0, (int)((c - lo) * scale)));
                return k <= best_bin;
            });
            mid = (size_t)(it - order.begin());
        } else {
            mid = b + n / 2; // coincident centres: split by count
        }
        if (mid == b || mid == e) mid = b + n / 2;

        // reserve the sibling pair before descending: parents precede children (depth-first order)
        const uint32_t pair = (uint32_t)(out->nodes.size() / 2);
        out->nodes.resize(out->nodes.size() + 4);
        out->max_depth = std::max(out->max_depth, depth + 1);
        Box lb, rb;
        const uint32_t lref = build(b, mid, depth + 1, &lb);
        const uint32_t rref = build(mid, e, depth + 1, &rb);
        write_node(pair, lb, lref);
        write_node(pair + 1, rb, rref);
        return pair;
    }

    void write_node(uint32_t i, const Box &bx, uint32_t ref) {
        F4 a, c;
        a.x = bx.lo[0], a.y = bx.lo[1], a.z = bx.lo[2];
        c.x = bx.hi[0], c.y = bx.hi[1], c.z = bx.hi[2], c.w = 0;
        memcpy(&a.w, &ref, 4);
        out->nodes[2 * (size_t)i] = a, out->nodes[2 * (size_t)i + 1] = c;
    }

    uint32_t make_leaf(size_t b, size_t e) {
        const uint32_t first = (uint32_t)out->sph.size();
        // keep object order inside a leaf: ties inside one leaf then resolve without a meta load
        std::sort(order.begin() + b, order.begin() + e);
        for (size_t i = b; i < e; i++) {
            const rt_sphere &s = sph[order[i]];
            F4 f;
            f.x = s.cx, f.y = s.cy, f.z = s.cz, f.w = s.r;
            out->sph.push_back(f);
            I2 m;
            m.x = (int32_t)order[i], m.y = (int32_t)s.material;
            out->meta.push_back(m);
        }
        return RT_LEAF | (first << 3) | (uint32_t)(e - b - 1);
    }
};

} // namespace

void compute_scene_center(const rt_sphere *spheres, uint64_t n, double m[3], double *extent90) {
    m[0] = m[1] = m[2] = 0, *extent90 = 0;
    if (n == 0) return;
    std::vector<float> v(n);
    for (int k = 0; k < 3; k++) {
        for (uint64_t i = 0; i < n; i++) v[i] = k == 0 ? spheres[i].cx : k == 1 ? spheres[i].cy : spheres[i].cz;
        std::nth_element(v.begin(), v.begin() + n / 2, v.end());
        m[k] = v[n / 2];
    }
    for (uint64_t i = 0; i < n; i++) {
        double dx = spheres[i].cx - m[0], dy = spheres[i].cy - m[1], dz = spheres[i].cz - m[2];
        v[i] = (float)(std::sqrt(dx * dx + dy * dy + dz * dz) + std::fabs((double)spheres[i].r));
    }
    size_t k90 = (size_t)((n - 1) * 0.9);
    std::nth_element(v.begin(), v.begin() + k90, v.end());
    *extent90 = v[k90];
}

// Padded box of one sphere for ray origins within origin_radius of m.
static Box padded_box(const rt_sphere &s, const double m[3], double origin_radius, float *pad_out) {
    double r = std::fabs((double)s.r);
    double dx = s.cx - m[0], dy = s.cy - m[1], dz = s.cz - m[2];
    double D = origin_radius + std::sqrt(dx * dx + dy * dy + dz * dz);
    double cmax = std::max(std::fabs((double)s.cx), std::max(std::fabs((double)s.cy), std::fabs((double)s.cz)));
    double mmax = std::max(std::fabs(m[0]), std::max(std::fabs(m[1]), std::fabs(m[2])));
    double pad = K_DISC * U * D * D / (2.0 * std::max(r, 1e-30)) + 8.0 * U * (D + mmax + cmax + r);
    pad = std::min(pad, D + r); // a box as large as the whole origin region is always enough
    Box b;
    const float c[3] = {s.cx, s.cy, s.cz};
    for (int k = 0; k < 3; k++) {
        b.lo[k] = std::nextafter((float)((double)c[k] - r - pad), -std::numeric_limits<float>::infinity());
        b.hi[k] = std::nextafter((float)((double)c[k] + r + pad), std::numeric_limits<float>::infinity());
    }
    *pad_out = (float)pad;
    return b;
}

void build_flat_bvh(const rt_sphere *spheres, uint64_t n, float origin_radius, int max_leaf, FlatBvh *out) {
    out->nodes.clear(), out->sph.clear(), out->meta.clear();
    out->root_ref = RT_REF_NONE, out->max_depth = 0;
    out->pad_min = out->pad_max = 0;
    if (n == 0) return;
    max_leaf = std::max(1, std::min(max_leaf, RT_MAX_LEAF));
    Builder b;
    b.sph = spheres, b.out = out, b.max_leaf = max_leaf;
    if (const char *e = getenv("RT_B200_BVH_BINS")) b.nbins = std::max(2, std::min(64, atoi(e)));
    if (const char *e = getenv("RT_B200_BVH_CTRAV")) b.c_trav = atof(e);
    b.boxes.resize(n), b.order.resize(n);
    double m[3], ext;
    compute_scene_center(spheres, n, m, &ext);
    float pmin = std::numeric_limits<float>::infinity(), pmax = 0;
    for (uint64_t i = 0; i < n; i++) {
        float pad;
        b.boxes[i] = padded_box(spheres[i], m, origin_radius, &pad);
        pmin = std::min(pmin, pad), pmax = std::max(pmax, pad);
        b.order[i] = (uint32_t)i;
    }
    out->pad_min = pmin, out->pad_max = pmax;
    out->nodes.reserve(4 * n);
    out->sph.reserve(n), out->meta.reserve(n);
    Box root_box;
    out->root_ref = b.build(0, n, 0, &root_box);
}

// Recompute every box for a larger origin_radius, topology unchanged.  Nodes are in pre-order
// (children after parents), so one reverse sweep rebuilds parents from children.
void refit_flat_bvh(const rt_sphere *spheres, uint64_t n, float origin_radius, FlatBvh *bvh) {
    if (n == 0 || bvh->root_ref == RT_REF_NONE) return;
    double m[3], ext;
    compute_scene_center(spheres, n, m, &ext);
    float pmin = std::numeric_limits<float>::infinity(), pmax = 0;
    const size_t n_nodes = bvh->nodes.size() / 2;
    for (size_t ii = n_nodes; ii-- > 0;) {
        uint32_t ref;
        memcpy(&ref, &bvh->nodes[2 * ii].w, 4);
        Box bx;
        bx.reset();
        if (ref & RT_LEAF) {
            uint32_t first = (ref & ~RT_LEAF) >> 3, count = (ref & 7u) + 1;
            for (uint32_t s = first; s < first + count; s++) {
                float pad;
                bx.grow(padded_box(spheres[bvh->meta[s].x], m, origin_radius, &pad));
                pmin = std::min(pmin, pad), pmax = std::max(pmax, pad);
            }
        } else {
            for (int c = 0; c < 2; c++) {
                const F4 &lo = bvh->nodes[2 * ((size_t)ref + c)], &hi = bvh->nodes[2 * ((size_t)ref + c) + 1];
                Box cbx;
                cbx.lo[0] = lo.x, cbx.lo[1] = lo.y, cbx.lo[2] = lo.z;
                cbx.hi[0] = hi.x, cbx.hi[1] = hi.y, cbx.hi[2] = hi.z;
                bx.grow(cbx);
            }
        }
        F4 &lo = bvh->nodes[2 * ii], &hi = bvh->nodes[2 * ii + 1];
        lo.x = bx.lo[0], lo.y = bx.lo[1], lo.z = bx.lo[2];
        hi.x = bx.hi[0], hi.y = bx.hi[1], hi.z = bx.hi[2];
    }
    bvh->pad_min = pmin, bvh->pad_max = pmax;
}

// Fold textures into 32-byte device material records (layout in rt_shade.h).
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

