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

struct Builder {
    const ScenePrims *prims;
    std::vector<Box> boxes;      // padded per-primitive boxes
    std::vector<float> cent;     // 3 per primitive
    std::vector<uint32_t> order; // permutation being partitioned
    FlatBvh *out;
    int max_leaf;

    static const int NBINS = 64; // upper bound; `nbins` bins are used
    int nbins_max = 64; // measured on C2: 16 -> 64 bins = 3 % fewer box tests per ray, +2.5 % Msamples/s
    double c_trav = 1.2; // cost of visiting a node pair, in primitive tests

    // returns the ref of the subtree over order[b, e), writes its box
    uint32_t build(size_t b, size_t e, uint32_t depth, Box *box_out) {
        Box bounds, cb;
        bounds.reset(), cb.reset();
        bool any_quad = false, any_sphere = false;
        for (size_t i = b; i < e; i++) {
            const uint32_t g = order[i];
            bounds.grow(boxes[g]);
            (is_quad(*prims, g) ? any_quad : any_sphere) = true;
            for (int k = 0; k < 3; k++)
                cb.lo[k] = std::min(cb.lo[k], cent[3 * g + k]), cb.hi[k] = std::max(cb.hi[k], cent[3 * g + k]);
        }
        *box_out = bounds;
        const size_t n = e - b;
        const int nbins = (int)std::min<size_t>((size_t)nbins_max, std::max<size_t>(4, n)); // no more bins than primitives
        // SAH over centroid bins per axis; cost unit = one primitive test
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
                    const uint32_t g = order[i];
                    int k = std::min(nbins - 1, std::max(0, (int)((cent[3 * g + axis] - lo) * scale)));
                    bin_box[k].grow(boxes[g]);
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
        // a leaf holds primitives of one kind; quads are one per leaf (80-byte records)
        const bool can_leaf = !(any_quad && any_sphere) && n <= (size_t)(any_quad ? 1 : max_leaf);
        if (can_leaf && (best_axis < 0 || (double)n <= split_cost)) return make_leaf(b, e);

        size_t mid;
        if (best_axis >= 0) {
            float lo = cb.lo[best_axis], ext = cb.hi[best_axis] - cb.lo[best_axis];
            float scale = (float)nbins / ext;
            auto it = std::partition(order.begin() + b, order.begin() + e, [&](uint32_t g) {
                int k = std::min(nbins - 1, std::max(0, (int)((cent[3 * g + best_axis] - lo) * scale)));
                return k <= best_bin;
            });
            mid = (size_t)(it - order.begin());
        } else if (any_quad && any_sphere) {
            auto it = std::partition(order.begin() + b, order.begin() + e, [&](uint32_t g) { return !is_quad(*prims, g); });
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
        const ScenePrims &p = *prims;
        if (is_quad(p, order[b])) {
            const uint32_t first = (uint32_t)out->quad_prim.size();
            for (size_t i = b; i < e; i++) {
                const uint32_t qi = order[i] - (uint32_t)p.spheres.size();
                out->quad.resize(out->quad.size() + RT_QUAD_F4);
                write_quad(p.quads[qi], p.quad_ids[qi], &out->quad[(size_t)RT_QUAD_F4 * out->quad_prim.size()]);
                out->quad_prim.push_back(qi);
            }
            return RT_LEAF | RT_LEAF_QUAD | (first << 3) | (uint32_t)(e - b - 1);
        }
        const uint32_t first = (uint32_t)out->sph.size();
        // ascending object ID inside a leaf
        std::sort(order.begin() + b, order.begin() + e, [&](uint32_t x, uint32_t y) { return p.sphere_ids[x] < p.sphere_ids[y]; });
        for (size_t i = b; i < e; i++) {
            const rt_sphere &s = p.spheres[order[i]];
            F4 f;
            f.x = s.cx, f.y = s.cy, f.z = s.cz, f.w = s.r;
            out->sph.push_back(f);
            I2 m;
            m.x = (int32_t)p.sphere_ids[order[i]], m.y = (int32_t)s.material;
            out->meta.push_back(m);
            out->sph_prim.push_back(order[i]);
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

// (min, max) boxes -> the (centre, half-extent) form the traversal kernels read.  The centre is rounded
// to float; the half-extent is taken from that rounded centre, rounded up, and widened by a few ulp of
// the magnitudes involved for the one extra rounding the centre form has in the slab test
// (m = (c - o) * inv is rounded before -+ h*|inv| is added).  Culling only has to be conservative.
static void make_device_nodes(FlatBvh *bvh) {
    const double U = 5.960464477539063e-08; // 2^-24
    bvh->dev_nodes.resize(bvh->nodes.size());
    for (size_t i = 0; i + 1 < bvh->nodes.size(); i += 2) {
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
}

void build_flat_bvh(const ScenePrims &prims, float origin_radius, int max_leaf, FlatBvh *out) {
    *out = FlatBvh();
    const size_t n = prims.size();
    if (n == 0) return;
    max_leaf = std::max(1, std::min(max_leaf, RT_MAX_LEAF));
    Builder b;
    b.prims = &prims, b.out = out, b.max_leaf = max_leaf;
    if (const char *e = getenv("RT_B200_BVH_BINS")) b.nbins_max = std::max(2, std::min(64, atoi(e)));
    if (const char *e = getenv("RT_B200_BVH_CTRAV")) b.c_trav = atof(e);
    b.boxes.resize(n), b.order.resize(n), b.cent.resize(3 * n);
    double m[3], ext;
    compute_scene_center(prims, m, &ext);
    float pmin = POS_INF, pmax = 0;
    for (size_t g = 0; g < n; g++) {
        float pad;
        b.boxes[g] = padded_box(prims, (uint32_t)g, m, origin_radius, &pad);
        pmin = std::min(pmin, pad), pmax = std::max(pmax, pad);
        double c[3], e;
        prim_center(prims, (uint32_t)g, c, &e);
        b.cent[3 * g] = (float)c[0], b.cent[3 * g + 1] = (float)c[1], b.cent[3 * g + 2] = (float)c[2];
        b.order[g] = (uint32_t)g;
    }
    out->pad_min = pmin, out->pad_max = pmax;
    out->nodes.reserve(4 * n);
    out->sph.reserve(prims.spheres.size()), out->meta.reserve(prims.spheres.size());
    Box root_box;
    out->root_ref = b.build(0, n, 0, &root_box);
    make_device_nodes(out);
}

// Recompute every box for a larger origin_radius, topology unchanged.  Nodes are in pre-order
// (children after parents), so one reverse sweep rebuilds parents from children.
void refit_flat_bvh(const ScenePrims &prims, float origin_radius, FlatBvh *bvh) {
    if (prims.size() == 0 || bvh->root_ref == RT_REF_NONE) return;
    double m[3], ext;
    compute_scene_center(prims, m, &ext);
    float pmin = POS_INF, pmax = 0;
    const size_t n_nodes = bvh->nodes.size() / 2;
    for (size_t ii = n_nodes; ii-- > 0;) {
        uint32_t ref;
        memcpy(&ref, &bvh->nodes[2 * ii].w, 4);
        Box bx;
        bx.reset();
        if (ref & RT_LEAF) {
            const bool quad = (ref & RT_LEAF_QUAD) != 0;
            uint32_t first = (ref & RT_LEAF_SLOT_MASK) >> 3, count = (ref & 7u) + 1;
            for (uint32_t s = first; s < first + count; s++) {
                float pad;
                const uint32_t g = quad ? (uint32_t)prims.spheres.size() + bvh->quad_prim[s] : bvh->sph_prim[s];
                bx.grow(padded_box(prims, g, m, origin_radius, &pad));
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

