// oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement of raytracer-go's per-pixel / per-sample path-tracing loop, written from
// the reference's source (citations `file:line` are relative to the reference checkout).  It is
// the checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / `--impl reference` legs may load it.  Nothing under raytracer_go_b200/ links,
// imports or calls it, and librt_b200.so has no CPU fallback.
//
// PARITY PIN STATUS — "parity unpinned" by the reference: the reference ships no tests, golden
// vectors or fixtures and Go is not installed here, so this restatement is
// pinned only by the hand-derived known-answer tests of SURVEY.md §4 (tests/test_oracle_kat.py)
// and by being an independent second implementation next to the device code.  What pins it wherever
// Go exists is committed: raytracer_go_b200/go/parity_dump_test.go evaluates tests/golden/pin_inputs.json with
// the reference's own functions, scripts/pin_from_go.sh runs it, tests/test_pin_from_go.py compares (the fed
// entry points orc_scatter_fed / orc_get_color_fed / orc_get_ray_fed below exist for that comparison).
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  With
// contraction off, x86-64 SSE float32 + - * / sqrt round exactly like Go on amd64 (go.mod:3,
// GOAMD64=v1 has no FMA), so every float32 expression below keeps the reference's operation
// order.  The RNG is NOT the reference's (math/rand, clock-seeded, camera.go:170): it is a
// counter-based Philox4x32-7 keyed by (seed; pixel, sample, block) — the published Random123
// algorithm (Salmon et al., SC'11) — consumed in whole blocks (layout at `struct Rng`) so that
// oracle and device draw identical values.

#include "../include/rt_b200.h"

#include <atomic>
#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// vec3.go
// ---------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
// vec3.go:43-53
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
// vec3.go:55-65
inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
// vec3.go:67-77
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
// vec3.go:91-101
inline V3 scale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
// vec3.go:115-117  (x*x + y*y) + z*z, left to right
inline float lensq(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
// vec3.go:137-139
inline float dot(V3 l, V3 r) { return l.x * r.x + l.y * r.y + l.z * r.z; }
// float32(math.Sqrt(float64(x))) — vec3.go:105, hittables.go:108
inline float sqrt32(float x) { return (float)std::sqrt((double)x); }
// vec3.go:103-107: scale by the reciprocal of the length
inline V3 unit(V3 a) {
    float l = sqrt32(lensq(a));
    return scale(a, 1 / l);
}
// vec3.go:129-135
inline V3 cross(V3 l, V3 r) {
    return v3(l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x);
}
// math.Min / math.Max on float64 (NaN-propagating), as used by math.go:38-44, materials.go:99
inline double go_min(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return std::numeric_limits<double>::quiet_NaN();
    if (a == 0 && b == 0) return std::signbit(a) ? a : b;
    return a < b ? a : b;
}
inline double go_max(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return std::numeric_limits<double>::quiet_NaN();
    if (a == 0 && b == 0) return std::signbit(a) ? b : a;
    return a > b ? a : b;
}
inline float minf32(float a, float b) { return (float)go_min((double)a, (double)b); } // math.go:38
inline float maxf32(float a, float b) { return (float)go_max((double)a, (double)b); } // math.go:42
// math.go:20-28
inline float clampf(float lo, float hi, float v) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
// vec3.go:168-172
inline bool near_zero(V3 v) {
    const float eps = 1e-8f;
    return (float)std::fabs((double)v.x) < eps && (float)std::fabs((double)v.y) < eps &&
           (float)std::fabs((double)v.z) < eps;
}
// vec3.go:212-214
inline V3 reflect(V3 v, V3 n) { return sub(v, scale(n, 2 * dot(v, n))); }
// vec3.go:216-221
inline V3 refract(V3 uv, V3 n, float eta) {
    float cos_theta = dot(scale(uv, -1), n);
    V3 perp = scale(add(uv, scale(n, cos_theta)), eta);
    float k = (float)std::sqrt(std::fabs((double)(1.0f - lensq(perp))));
    V3 par = scale(n, -1 * k);
    return add(par, perp);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-R (Random123).  Stream of one path: counter (pixel, sample, block, 0),
// key (seed_lo, seed_hi); the four words of a block are consumed in order.  The render streams use
// g_philox_rounds = 7 (the fewest rounds Random123 reports as Crush-resistant; csrc/rt_rng.h says why);
// orc_set_philox_rounds(10) switches to Random123's default (the committed converged frame was drawn with it).
// ---------------------------------------------------------------------------------------------
int g_philox_rounds = 7;
inline void philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], int rounds) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

// Stream layout: every consumer takes WHOLE blocks, block b = philox(ctr = (pixel, sample, b, 0)):
//   GetRay: one block = (dx, dy, disk.x, disk.y), then one more block (two candidate pairs) per
//   rejected disk pair; unit-sphere sampler: one block per trial (x, y, z, unused);
//   Dielectric.Scatter: one block, first word.
struct Block {
    float u[4];
};
struct Rng {
    uint32_t ctr[4];
    uint32_t key[2];
    Rng(uint64_t seed, uint32_t pixel, uint32_t sample) {
        ctr[0] = pixel, ctr[1] = sample, ctr[2] = 0, ctr[3] = 0;
        key[0] = (uint32_t)seed, key[1] = (uint32_t)(seed >> 32);
    }
    // Reference pin (tests/test_pin_from_go.py): instead of Philox, hand out the uniforms of a caller-supplied list,
    // four per block, so that the real Go code (fed the same values through a stub rand.Source) and this restatement
    // see identical draws.  An exhausted list repeats its last value.
    const float *feed = nullptr;
    int64_t feed_n = 0, feed_pos = 0;
    Rng(const float *values, int64_t n) : feed(values), feed_n(n) {
        ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0, key[0] = key[1] = 0;
    }
    // rand.Float32() is uniform on [0,1) (camera.go:290); 24 random mantissa bits.
    Block next() {
        if (feed) {
            Block b;
            for (int i = 0; i < 4; i++, feed_pos++) b.u[i] = feed_n ? feed[std::min(feed_pos, feed_n - 1)] : 0.0f;
            return b;
        }
        uint32_t w[4];
        philox4x32(ctr, key, w, g_philox_rounds);
        ctr[2]++;
        Block b;
        for (int i = 0; i < 4; i++) b.u[i] = (float)(w[i] >> 8) * (1.0f / 16777216.0f);
        return b;
    }
};

// math.go:30-32
inline float rand_range(float r, float lo, float hi) { return lo + r * (hi - lo); }
// vec3.go:182-190 (vec3.go:178-180 draws x, y, z in order)
inline V3 rand_unit(Rng &r) {
    for (;;) {
        Block b = r.next();
        V3 v = v3(rand_range(b.u[0], -1, 1), rand_range(b.u[1], -1, 1), rand_range(b.u[2], -1, 1));
        if (lensq(v) < 1.0f) return unit(v);
    }
}
// vec3.go:203-210: `first` is the pair left over in GetRay's block
inline V3 rand_in_unit_disk(Rng &r, float u2, float u3) {
    V3 v = v3(rand_range(u2, -1, 1), rand_range(u3, -1, 1), 0);
    if (lensq(v) < 1) return v;
    for (;;) {
        Block b = r.next();
        v = v3(rand_range(b.u[0], -1, 1), rand_range(b.u[1], -1, 1), 0);
        if (lensq(v) < 1) return v;
        v = v3(rand_range(b.u[2], -1, 1), rand_range(b.u[3], -1, 1), 0);
        if (lensq(v) < 1) return v;
    }
}

// ---------------------------------------------------------------------------------------------
// ray.go, bvh.go, hittables.go
// ---------------------------------------------------------------------------------------------
struct Ray {
    V3 origin, dir;
};
// ray.go:25-30: dir*t then + origin
inline V3 ray_at(const Ray &r, float t) { return add(scale(r.dir, t), r.origin); }

struct Interval {
    float min, max;
};
// bvh.go:18-20 with padding 0
inline bool interval_in(Interval i, float v) { return i.min < v && v < i.max; }

struct Aabb {
    Interval x, y, z;
};
// bvh.go:28-34
inline Aabb aabb_from_points(V3 p1, V3 p2) {
    return Aabb{{minf32(p1.x, p2.x), maxf32(p1.x, p2.x)},
                {minf32(p1.y, p2.y), maxf32(p1.y, p2.y)},
                {minf32(p1.z, p2.z), maxf32(p1.z, p2.z)}};
}
// bvh.go:44-50
inline Aabb aabb_union(const Aabb &a, const Aabb &b) {
    return Aabb{{minf32(a.x.min, b.x.min), maxf32(a.x.max, b.x.max)},
                {minf32(a.y.min, b.y.min), maxf32(a.y.max, b.y.max)},
                {minf32(a.z.min, b.z.min), maxf32(a.z.max, b.z.max)}};
}
// bvh.go:84-102
inline bool in_boundary(float dir, float origin, float amin, float amax, Interval *rt) {
    float inv_d = 1 / dir;
    float t0 = (amin - origin) * inv_d;
    float t1 = (amax - origin) * inv_d;
    if (inv_d < 0) std::swap(t0, t1);
    if (t0 > rt->min) rt->min = t0;
    if (t1 < rt->max) rt->max = t1;
    return rt->min < rt->max;
}
// bvh.go:52-61 (the interval is a by-value copy: the caller's is unchanged)
inline bool aabb_hit(const Aabb &a, const Ray &r, Interval rt) {
    if (in_boundary(r.dir.x, r.origin.x, a.x.min, a.x.max, &rt))
        if (in_boundary(r.dir.y, r.origin.y, a.y.min, a.y.max, &rt))
            if (in_boundary(r.dir.z, r.origin.z, a.z.min, a.z.max, &rt)) return true;
    return false;
}

// hittables.go:12-20
struct HitInfo {
    V3 point, normal;
    float t, u, v;
    int32_t object; // index into World.hittables (stands in for the Material interface value)
    bool front_face;
};
// hittables.go:22-37
inline HitInfo new_hit_info(float t, float u, float v, V3 ray_dir, V3 point, V3 outward,
                            int32_t object) {
    bool front = dot(ray_dir, outward) < 0;
    if (!front) outward = scale(outward, -1);
    return HitInfo{point, outward, t, u, v, object, front};
}

struct Counters {
    uint64_t rays = 0, box_tests = 0, sphere_tests = 0, hits = 0, samples = 0;
};
thread_local Counters g_cnt;

const float PI_F32 = (float)M_PI; // math.go:48

// hittables.go:96-132
inline bool sphere_hit(const rt_sphere &s, int32_t object, const Ray &r, Interval rt, HitInfo *out) {
    g_cnt.sphere_tests++;
    V3 center = v3(s.cx, s.cy, s.cz);
    V3 a_sub_c = sub(r.origin, center);
    float a = lensq(r.dir);
    float half_b = dot(r.dir, a_sub_c);
    float c = lensq(a_sub_c) - s.r * s.r;
    float disc = (half_b * half_b - a * c);
    if (disc < 0) return false;
    float sqt = sqrt32(disc);
    float t;
    float root = (-half_b - sqt) / a;
    if (interval_in(rt, root)) {
        t = root;
    } else {
        root = (-half_b + sqt) / a;
        if (interval_in(rt, root)) t = root;
        else return false;
    }
    V3 point = ray_at(r, t);
    V3 norm = unit(scale(sub(point, center), s.r));
    float theta = (float)std::acos(-(double)norm.y);
    float phi = (float)(std::atan2(-(double)norm.z, (double)norm.x) + M_PI);
    float u = (phi + 5 * PI_F32 / 12) / (2 * PI_F32);
    float v = theta / (PI_F32);
    *out = new_hit_info(t, u, v, r.dir, point, norm, object);
    return true;
}
// hittables.go:85-94
inline Aabb sphere_bounds(const rt_sphere &s) {
    V3 c = v3(s.cx, s.cy, s.cz);
    V3 rvec = v3(s.r, s.r, s.r);
    return aabb_from_points(add(c, scale(rvec, -1)), add(c, rvec));
}

// hittables.go:138-147 plus the fields NewQuad derives (hittables.go:149-165)
struct Quad {
    V3 Q, u, v, w, normal;
    float D;
    uint32_t material;
    Aabb bbox;
};

// bvh.go:63-82
inline Aabb padded_aabb(Aabb a) {
    const float eps = 0.0001f;
    if (a.x.max - a.x.min < eps) a.x.min -= eps, a.x.max += eps;
    if (a.y.max - a.y.min < eps) a.y.min -= eps, a.y.max += eps;
    if (a.z.max - a.z.min < eps) a.z.min -= eps, a.z.max += eps;
    return a;
}

// hittables.go:149-165
inline Quad new_quad(const rt_quad &q) {
    Quad r;
    r.Q = v3(q.q[0], q.q[1], q.q[2]), r.u = v3(q.u[0], q.u[1], q.u[2]), r.v = v3(q.v[0], q.v[1], q.v[2]);
    V3 n = cross(r.u, r.v);
    r.normal = unit(n);
    r.D = dot(r.normal, r.Q);
    r.w = scale(n, 1 / dot(n, n));
    r.material = q.material;
    r.bbox = padded_aabb(aabb_from_points(r.Q, add(add(r.Q, r.u), r.v)));
    return r;
}

// hittables.go:167-194
inline bool quad_hit(const Quad &q, int32_t object, const Ray &r, Interval rt, HitInfo *out) {
    g_cnt.sphere_tests++;
    float denom = dot(r.dir, q.normal);
    if (std::fabs((double)denom) < 1e-8) return false;
    float t = (q.D - dot(q.normal, r.origin)) / denom;
    if (!interval_in(rt, t)) return false;
    V3 intersection = ray_at(r, t);
    V3 plane_hit = sub(intersection, q.Q);
    float alpha = dot(q.w, cross(plane_hit, q.v));
    float beta = dot(q.w, cross(q.u, plane_hit));
    if (alpha < 0 || 1 < alpha || beta < 0 || 1 < beta) return false; // InPlane, hittables.go:192-194
    *out = new_hit_info(t, alpha, beta, r.dir, intersection, q.normal, object);
    return true;
}

struct Prim {
    int kind; // 0 sphere, 1 quad
    uint32_t index;
};

struct Scene {
    std::vector<rt_sphere> spheres;
    std::vector<Quad> quads;
    std::vector<Prim> prims; // World.hittables: position = object ID (hittables.go:48-53)
    std::vector<rt_material> materials;
    std::vector<rt_texture> textures;
    struct Img {
        int w, h;
        std::vector<uint16_t> px;
    };
    std::vector<Img> images;
    std::vector<rt_perlin> perlins;
    uint32_t material_of(int32_t object) const {
        const Prim &p = prims[object];
        return p.kind == 0 ? spheres[p.index].material : quads[p.index].material;
    }
    Aabb bounds_of(int32_t object) const {
        const Prim &p = prims[object];
        return p.kind == 0 ? sphere_bounds(spheres[p.index]) : quads[p.index].bbox;
    }
    bool hit_prim(int32_t object, const Ray &r, Interval rt, HitInfo *out) const {
        const Prim &p = prims[object];
        return p.kind == 0 ? sphere_hit(spheres[p.index], object, r, rt, out) : quad_hit(quads[p.index], object, r, rt, out);
    }
};

bool load_scene(const rt_scene_desc *d, Scene *s) {
    if (!d) return false;
    s->spheres.assign(d->spheres, d->spheres + d->n_spheres);
    for (uint64_t i = 0; i < d->n_quads; i++) s->quads.push_back(new_quad(d->quads[i]));
    const size_t n = s->spheres.size() + s->quads.size();
    s->prims.assign(n, Prim{-1, 0});
    for (size_t i = 0; i < s->spheres.size(); i++) {
        size_t id = d->sphere_ids ? d->sphere_ids[i] : i;
        if (id >= n || s->prims[id].kind != -1) return false;
        s->prims[id] = Prim{0, (uint32_t)i};
    }
    for (size_t i = 0; i < s->quads.size(); i++) {
        size_t id = d->quad_ids ? d->quad_ids[i] : s->spheres.size() + i;
        if (id >= n || s->prims[id].kind != -1) return false;
        s->prims[id] = Prim{1, (uint32_t)i};
    }
    s->materials.assign(d->materials, d->materials + d->n_materials);
    s->textures.assign(d->textures, d->textures + d->n_textures);
    for (uint32_t i = 0; i < d->n_images; i++) {
        Scene::Img im;
        im.w = d->images[i].w, im.h = d->images[i].h;
        size_t np = (size_t)std::max(im.w, 0) * (size_t)std::max(im.h, 0) * 3;
        im.px.assign(d->images[i].rgb16, d->images[i].rgb16 + np);
        s->images.push_back(std::move(im));
    }
    if (d->n_perlins) s->perlins.assign(d->perlins, d->perlins + d->n_perlins);
    for (size_t i = 0; i < n; i++)
        if (s->material_of((int32_t)i) >= s->materials.size()) return false;
    for (auto &m : s->materials)
        if ((m.kind == RT_MAT_LAMBERTIAN || m.kind == RT_MAT_DIFFUSE_LIGHT) &&
            m.texture >= s->textures.size())
            return false;
    for (auto &t : s->textures)
        if ((t.kind == RT_TEX_IMAGE && t.image >= s->images.size()) ||
            (t.kind == RT_TEX_NOISE && t.image >= s->perlins.size()))
            return false;
    return true;
}

// hittables.go:55-72: insertion-ordered brute force; the first object wins exact ties.
inline bool world_hit(const Scene &sc, const Ray &r, Interval rt, HitInfo *out) {
    bool hit_any = false;
    float closest = rt.max;
    for (size_t i = 0; i < sc.prims.size(); i++) {
        HitInfo hi{};
        if (sc.hit_prim((int32_t)i, r, Interval{rt.min, closest}, &hi)) {
            hit_any = true;
            *out = hi;
            closest = hi.t;
        }
    }
    return hit_any;
}

// bvh.go:132-136.  A child is either an inner node (index >= 0 into nodes) or a primitive.
struct RefBvhNode {
    int32_t left_node, right_node; // -1 when the child is a primitive
    int32_t left_prim, right_prim;
    Aabb box;
};
struct RefBvh {
    std::vector<RefBvhNode> nodes;
    int32_t root = -1;
};

// A small deterministic stand-in for the global rand.Intn(3) of bvh.go:147 (its stream is not
// part of any contract: the reference's tree is different on every run).
struct SplitMix {
    uint64_t s;
    uint32_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return (uint32_t)((z ^ (z >> 31)) >> 32);
    }
};

// bvh.go:187-218: "descending by bbox min": compare(h1,h2) = sign(h2.min - h1.min)
inline int ref_compare(const Aabb &b1, const Aabb &b2, int axis) {
    float m1 = axis == 0 ? b1.x.min : axis == 1 ? b1.y.min : b1.z.min;
    float m2 = axis == 0 ? b2.x.min : axis == 1 ? b2.y.min : b2.z.min;
    float diff = m2 - m1;
    if (diff > 0) return 1;
    if (diff < 0) return -1;
    return 0;
}

// bvh.go:142-185
int32_t ref_bvh_build(const Scene &sc, RefBvh &bvh, std::vector<int32_t> prims, SplitMix &rng) {
    int axis = (int)(rng.next() % 3);
    RefBvhNode node{};
    node.left_node = node.right_node = -1;
    node.left_prim = node.right_prim = -1;
    Aabb lb, rb;
    if (prims.size() == 1) {
        node.left_prim = node.right_prim = prims[0];
        lb = rb = sc.bounds_of(prims[0]);
    } else if (prims.size() == 2) {
        Aabb b0 = sc.bounds_of(prims[0]), b1 = sc.bounds_of(prims[1]);
        if (ref_compare(b0, b1, axis) > 0) {
            node.left_prim = prims[1], node.right_prim = prims[0];
            lb = b1, rb = b0;
        } else {
            node.left_prim = prims[0], node.right_prim = prims[1];
            lb = b0, rb = b1;
        }
    } else {
        std::stable_sort(prims.begin(), prims.end(), [&](int32_t p, int32_t q) {
            return ref_compare(sc.bounds_of(p), sc.bounds_of(q), axis) < 0;
        });
        size_t mid = prims.size() / 2;
        std::vector<int32_t> l(prims.begin(), prims.begin() + mid), r(prims.begin() + mid, prims.end());
        node.left_node = ref_bvh_build(sc, bvh, std::move(l), rng);
        node.right_node = ref_bvh_build(sc, bvh, std::move(r), rng);
        lb = bvh.nodes[node.left_node].box;
        rb = bvh.nodes[node.right_node].box;
    }
    node.box = aabb_union(lb, rb); // bvh.go:182
    bvh.nodes.push_back(node);
    return (int32_t)bvh.nodes.size() - 1;
}

// bvh.go:220-249
bool ref_bvh_hit(const Scene &sc, const RefBvh &bvh, int32_t n, const Ray &r, Interval rt, HitInfo *out) {
    const RefBvhNode &b = bvh.nodes[n];
    g_cnt.box_tests++;
    if (!aabb_hit(b.box, r, rt)) return false;
    HitInfo lh{}, rh{};
    bool hit_left = b.left_node >= 0 ? ref_bvh_hit(sc, bvh, b.left_node, r, rt, &lh)
                                     : sc.hit_prim(b.left_prim, r, rt, &lh);
    Interval right_rt = rt;
    if (hit_left) right_rt.max = lh.t;
    bool hit_right = b.right_node >= 0 ? ref_bvh_hit(sc, bvh, b.right_node, r, right_rt, &rh)
                                       : sc.hit_prim(b.right_prim, r, right_rt, &rh);
    if (hit_left && hit_right) {
        *out = lh.t < rh.t ? lh : rh;
        return true;
    }
    if (hit_right) {
        *out = rh;
        return true;
    }
    if (hit_left) {
        *out = lh;
        return true;
    }
    return false;
}

enum { MODE_LINEAR = 0, MODE_REF_BVH = 1 };
enum { ORDER_RECURSIVE = 0, ORDER_ITERATIVE = 1 };

struct World {
    const Scene *sc;
    const RefBvh *bvh; // null => World.Hit
    bool hit(const Ray &r, Interval rt, HitInfo *out) const {
        if (bvh) {
            if (bvh->root < 0) return false;
            return ref_bvh_hit(*sc, *bvh, bvh->root, r, rt, out);
        }
        return world_hit(*sc, r, rt, out);
    }
};

// ---------------------------------------------------------------------------------------------
// materials.go
// ---------------------------------------------------------------------------------------------
// materials.go:127-137
inline V3 checker_texture(const rt_texture &t, V3 p) {
    float inv = 1 / t.scale;
    long long x = (long long)std::floor((double)(inv * p.x));
    long long y = (long long)std::floor((double)(inv * p.y));
    long long z = (long long)std::floor((double)(inv * p.z));
    if ((x + y + z) % 2 == 0) return v3(t.a[0], t.a[1], t.a[2]);
    return v3(t.b[0], t.b[1], t.b[2]);
}
// materials.go:175-193
inline V3 image_texture(const Scene &sc, const rt_texture &t, float u, float v) {
    const Scene::Img &im = sc.images[t.image];
    if (im.h <= 0) return v3(0, 1, 1);
    u = clampf(0, 1, u);
    v = 1 - clampf(0, 1, v);
    float fi = u * (float)im.w;
    float fj = v * (float)im.h;
    int i = (int)fi, j = (int)fj;
    // image.Image.At outside Bounds() returns the zero colour of the image's model
    if (i < 0 || i >= im.w || j < 0 || j >= im.h) return v3(t.oob[0], t.oob[1], t.oob[2]);
    const uint16_t *px = &im.px[((size_t)j * im.w + i) * 3];
    float col_scale = (float)(1.0 / 65535.0);
    return v3((float)px[0] * col_scale, (float)px[1] * col_scale, (float)px[2] * col_scale);
}
// math.go:58-60
inline float lerp1(float t, float x, float y) { return x * (1 - t) + y * t; }
// math.go:78-82
inline float bilerp(float tx, float ty, float c00, float c10, float c01, float c11) {
    float a = lerp1(tx, c00, c10);
    float b = lerp1(tx, c01, c11);
    return lerp1(ty, a, b);
}
// math.go:84-92
inline float trilerp(float tx, float ty, float tz, float c000, float c100, float c010, float c110, float c001,
                     float c101, float c011, float c111) {
    float e = bilerp(tx, ty, c000, c100, c010, c110);
    float f = bilerp(tx, ty, c001, c101, c011, c111);
    return lerp1(tz, e, f);
}
// materials.go:218-220
inline float smoothstep(float t) { return t * t * (3 - 2 * t); }
// materials.go:223-249
inline float perlin_noise(const rt_perlin &per, V3 p) {
    float xi = (float)std::floor((double)p.x), yi = (float)std::floor((double)p.y), zi = (float)std::floor((double)p.z);
    float tx = p.x - xi, ty = p.y - yi, tz = p.z - zi;
    int rx0 = (int)((long long)xi & 255), rx1 = (rx0 + 1) & 255;
    int ry0 = (int)((long long)yi & 255), ry1 = (ry0 + 1) & 255;
    int rz0 = (int)((long long)zi & 255), rz1 = (rz0 + 1) & 255;
    auto g = [&](int ix, int iy, int iz) {
        const float *q = per.vec[per.perm_x[ix] ^ per.perm_y[iy] ^ per.perm_z[iz]];
        return v3(q[0], q[1], q[2]);
    };
    float c000 = dot(g(rx0, ry0, rz0), v3(tx, ty, tz));
    float c001 = dot(g(rx0, ry0, rz1), v3(tx, ty, tz - 1));
    float c010 = dot(g(rx0, ry1, rz0), v3(tx, ty - 1, tz));
    float c011 = dot(g(rx0, ry1, rz1), v3(tx, ty - 1, tz - 1));
    float c100 = dot(g(rx1, ry0, rz0), v3(tx - 1, ty, tz));
    float c101 = dot(g(rx1, ry0, rz1), v3(tx - 1, ty, tz - 1));
    float c110 = dot(g(rx1, ry1, rz0), v3(tx - 1, ty - 1, tz));
    float c111 = dot(g(rx1, ry1, rz1), v3(tx - 1, ty - 1, tz - 1));
    return trilerp(smoothstep(tx), smoothstep(ty), smoothstep(tz), c000, c100, c010, c110, c001, c101, c011, c111);
}
// materials.go:251-262
inline float perlin_turb(const rt_perlin &per, V3 p, int depth) {
    float sum = 0, weight = 1.0f;
    for (int i = 0; i < depth; i++) {
        sum += weight * perlin_noise(per, p);
        weight *= 0.5f;
        p = scale(p, 2);
    }
    return (float)std::fabs((double)sum);
}
// materials.go:285-288
inline V3 noise_texture(const Scene &sc, const rt_texture &t, V3 point) {
    point = scale(point, t.scale);
    float s = 0.5f * (1 + (float)std::sin((double)(point.z + 10 * perlin_turb(sc.perlins[t.image], point, 7))));
    return scale(v3(1, 1, 1), s);
}

inline V3 texture_value(const Scene &sc, uint32_t tex, float u, float v, V3 p) {
    const rt_texture &t = sc.textures[tex];
    switch (t.kind) {
    case RT_TEX_NOISE: return noise_texture(sc, t, p);
    case RT_TEX_CHECKER: return checker_texture(t, p);
    case RT_TEX_IMAGE: return image_texture(sc, t, u, v);
    default: return v3(t.a[0], t.a[1], t.a[2]); // materials.go:155-157
    }
}
// materials.go:115-119
inline float reflectance(float cos_theta, float eta) {
    float r0 = (1.0f - eta) / (1.0f + eta);
    r0 *= r0;
    return r0 + (1 - r0) * (float)std::pow(1 - (double)cos_theta, 5);
}

struct Scatter {
    Ray ray;
    V3 attenuation;
};

// Material.Emit: materials.go:23,49,81 return zero; materials.go:311 returns the texture.
inline V3 material_emit(const Scene &sc, const rt_material &m, const HitInfo &hi) {
    if (m.kind == RT_MAT_DIFFUSE_LIGHT) return texture_value(sc, m.texture, hi.u, hi.v, hi.point);
    return v3(0, 0, 0);
}

inline bool material_scatter(const Scene &sc, const rt_material &m, const Ray &r, const HitInfo &hi,
                             Rng &rng, Scatter *out) {
    switch (m.kind) {
    case RT_MAT_LAMBERTIAN: { // materials.go:33-42
        V3 dir = add(hi.normal, rand_unit(rng));
        if (near_zero(dir)) dir = hi.normal;
        out->ray = Ray{hi.point, dir};
        out->attenuation = texture_value(sc, m.texture, hi.u, hi.v, hi.point);
        return true;
    }
    case RT_MAT_METAL: { // materials.go:60-75
        V3 unit_dir = unit(r.dir);
        V3 reflected = reflect(unit_dir, hi.normal);
        V3 fuzz = scale(rand_unit(rng), m.fuzz);
        V3 scattered = add(reflected, fuzz);
        if (dot(scattered, hi.normal) > 0) {
            out->ray = Ray{hi.point, scattered};
            out->attenuation = v3(m.albedo[0], m.albedo[1], m.albedo[2]);
            return true;
        }
        return false;
    }
    case RT_MAT_DIELECTRIC: { // materials.go:91-113
        float eta = m.ior;
        if (hi.front_face) eta = 1.0f / m.ior;
        V3 unit_dir = unit(r.dir);
        float cos_theta = (float)go_min((double)dot(scale(unit_dir, -1), hi.normal), 1.0);
        float sin_theta = (float)std::sqrt(1 - (double)(cos_theta * cos_theta));
        bool cannot_refract = sin_theta * eta > 1.0f;
        V3 direction;
        // `||` short-circuits: the uniform is drawn only when refraction is possible
        if (cannot_refract || reflectance(cos_theta, eta) > rng.next().u[0])
            direction = reflect(unit_dir, hi.normal);
        else
            direction = refract(unit_dir, hi.normal, eta);
        out->ray = Ray{hi.point, direction};
        out->attenuation = v3(1, 1, 1);
        return true;
    }
    default: // DiffuseLight.Scatter, materials.go:301-303
        return false;
    }
}

// ---------------------------------------------------------------------------------------------
// ray.go:32-54
// ---------------------------------------------------------------------------------------------
const float T_MIN = 0.001f;

V3 get_color_recursive(const World &w, const Ray &r, V3 background, int max_depth, Rng &rng) {
    if (max_depth <= 0) return v3(0, 0, 0);
    g_cnt.rays++;
    HitInfo hi{};
    if (w.hit(r, Interval{T_MIN, std::numeric_limits<float>::infinity()}, &hi)) {
        g_cnt.hits++;
        const rt_material &m = w.sc->materials[w.sc->material_of(hi.object)];
        V3 emitted = material_emit(*w.sc, m, hi);
        Scatter sc;
        if (!material_scatter(*w.sc, m, r, hi, rng, &sc)) return emitted;
        V3 scattered = mul(sc.attenuation, get_color_recursive(w, sc.ray, background, max_depth - 1, rng));
        return add(emitted, scattered);
    }
    return background;
}

// The same radiance with the recursion unrolled front to back (the order the device uses):
//   L = sum_i T_i * E_i  +  T_n * background,  T_0 = 1,  T_{i+1} = T_i * A_i.
// Differs from get_color_recursive only in float32 rounding order of the products.
V3 get_color_iterative(const World &w, Ray r, V3 background, int max_depth, Rng &rng) {
    V3 throughput = v3(1, 1, 1);
    V3 radiance = v3(0, 0, 0);
    for (int depth = 0; depth < max_depth; depth++) {
        g_cnt.rays++;
        HitInfo hi{};
        if (!w.hit(r, Interval{T_MIN, std::numeric_limits<float>::infinity()}, &hi))
            return add(radiance, mul(throughput, background));
        g_cnt.hits++;
        const rt_material &m = w.sc->materials[w.sc->material_of(hi.object)];
        V3 emitted = material_emit(*w.sc, m, hi);
        radiance = add(radiance, mul(throughput, emitted));
        Scatter sc;
        if (!material_scatter(*w.sc, m, r, hi, rng, &sc)) return radiance;
        throughput = mul(throughput, sc.attenuation);
        r = sc.ray;
    }
    return radiance;
}

// ---------------------------------------------------------------------------------------------
// camera.go
// ---------------------------------------------------------------------------------------------
// camera.go:128-166
void camera_init(const rt_camera_options &o, rt_camera *c) {
    V3 look_from = v3(o.look_from[0], o.look_from[1], o.look_from[2]);
    V3 look_at = v3(o.look_at[0], o.look_at[1], o.look_at[2]);
    V3 vup = v3(o.vup[0], o.vup[1], o.vup[2]);
    float image_width = (float)o.image_width; // camera.go:107
    V3 center = look_from;
    V3 dist = sub(look_from, look_at);
    float h = (float)std::tan((double)(o.fov_radians / 2.0f));
    float viewport_height = 2.0f * h * o.focus_dist;
    float image_height = (float)(std::floor((double)image_width) / (double)o.aspect_ratio);
    if (image_height < 1) image_height = 1;
    float viewport_width = viewport_height * (image_width / image_height);
    V3 w = unit(dist);
    V3 u = unit(cross(vup, w));
    V3 v = cross(w, u);
    V3 viewport_u = scale(u, viewport_width);
    V3 viewport_v = scale(v, -viewport_height);
    V3 pixel_du = scale(viewport_u, 1 / image_width);
    V3 pixel_dv = scale(viewport_v, 1 / image_height);
    V3 upper_left = center;
    upper_left = sub(upper_left, scale(w, o.focus_dist));
    upper_left = sub(upper_left, scale(viewport_u, 0.5f));
    upper_left = sub(upper_left, scale(viewport_v, 0.5f));
    V3 pixel00 = add(upper_left, scale(add(pixel_du, pixel_dv), 0.5f));
    float defocus_radius = o.focus_dist * (float)std::tan((double)(o.defocus_angle_radians / 2.0f));
    V3 disk_u = scale(u, defocus_radius);
    V3 disk_v = scale(v, defocus_radius);

    c->width = (int32_t)image_width;   // camera.go:181
    c->height = (int32_t)image_height; // camera.go:182
    c->spp = o.spp;
    c->max_depth = o.max_depth;
    auto put = [](float *d, V3 s) { d[0] = s.x, d[1] = s.y, d[2] = s.z; };
    put(c->center, center);
    put(c->pixel00, pixel00);
    put(c->pixel_du, pixel_du);
    put(c->pixel_dv, pixel_dv);
    put(c->defocus_u, disk_u);
    put(c->defocus_v, disk_v);
    c->defocus_angle = o.defocus_angle_radians;
    c->background[0] = o.background[0], c->background[1] = o.background[1], c->background[2] = o.background[2];
}

inline V3 ld3(const float *p) { return v3(p[0], p[1], p[2]); }

// camera.go:265-299
inline Ray get_ray(const rt_camera &c, Rng &rng, int i, int j) {
    V3 du_off = scale(ld3(c.pixel_du), (float)i);
    V3 dv_off = scale(ld3(c.pixel_dv), (float)j);
    V3 pixel_center = ld3(c.pixel00);
    pixel_center = add(pixel_center, du_off);
    pixel_center = add(pixel_center, dv_off);
    // sampleUnitSquare, camera.go:289-299
    Block b = rng.next();
    float dx = -0.5f + b.u[0];
    float dy = -0.5f + b.u[1];
    pixel_center = add(pixel_center, add(scale(ld3(c.pixel_du), dx), scale(ld3(c.pixel_dv), dy)));
    V3 disc = rand_in_unit_disk(rng, b.u[2], b.u[3]); // always drawn, camera.go:277
    V3 origin = ld3(c.center);
    if (c.defocus_angle > 0)
        origin = add(ld3(c.center), add(scale(ld3(c.defocus_u), disc.x), scale(ld3(c.defocus_v), disc.y)));
    V3 dir = sub(pixel_center, origin);
    return Ray{origin, dir};
}

// vec3.go:162-166, 145-152, and the int() of vec3.go:141-143
inline void encode_pixel(V3 mean, uint8_t *rgb) {
    float ch[3] = {mean.x, mean.y, mean.z};
    for (int k = 0; k < 3; k++) {
        float g = sqrt32(ch[k]);
        g = clampf(0, 1, g);
        g *= 255.999f;
        rgb[k] = std::isnan(g) ? 0 : (uint8_t)(int)g;
    }
}

} // namespace

// =============================================================================================
// C interface (ctypes from tests/ and bench.py only)
// =============================================================================================
extern "C" {

struct orc_stats {
    uint64_t samples, rays, box_tests, sphere_tests, hits;
    double seconds;
    int32_t threads;
    int32_t reserved;
};

int orc_camera_from_options(const rt_camera_options *o, rt_camera *out) {
    if (!o || !out) return -1;
    camera_init(*o, out);
    return 0;
}

void orc_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32(ctr, key, out, 10); }
void orc_philox4x32(const uint32_t *ctr, const uint32_t *key, int rounds, uint32_t *out) { philox4x32(ctr, key, out, rounds); }
// Rounds of the render streams (7 by default, as the device); returns the previous value.
int orc_set_philox_rounds(int rounds) {
    const int prev = g_philox_rounds;
    if (rounds >= 1 && rounds <= 16) g_philox_rounds = rounds;
    return prev;
}

// The uniforms of the first ceil(n/4) blocks of the stream of (seed, pixel, sample).
void orc_rng_floats(uint64_t seed, uint32_t pixel, uint32_t sample, int n, float *out) {
    Rng r(seed, pixel, sample);
    for (int i = 0; i < n; i += 4) {
        Block b = r.next();
        for (int k = 0; k < 4 && i + k < n; k++) out[i + k] = b.u[k];
    }
}

// World.Hit (mode 0) or the reference's BVH.Hit over a reference-style tree (mode 1).
int orc_trace(const rt_scene_desc *desc, int mode, uint64_t bvh_seed, const float *origins,
              const float *dirs, int64_t n, float tmin, float tmax, int32_t *id_out, float *t_out,
              int threads) {
    Scene sc;
    if (!load_scene(desc, &sc)) return -1;
    RefBvh bvh;
    if (mode == MODE_REF_BVH && !sc.prims.empty()) {
        std::vector<int32_t> prims(sc.prims.size());
        for (size_t i = 0; i < prims.size(); i++) prims[i] = (int32_t)i;
        SplitMix rng{bvh_seed};
        bvh.root = ref_bvh_build(sc, bvh, prims, rng);
    }
    World w{&sc, mode == MODE_REF_BVH ? &bvh : nullptr};
    if (threads < 1) threads = 1;
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            int64_t b = next.fetch_add(4096);
            if (b >= n) break;
            int64_t e = std::min(n, b + 4096);
            for (int64_t i = b; i < e; i++) {
                Ray r{ld3(origins + 3 * i), ld3(dirs + 3 * i)};
                HitInfo hi{};
                if (w.hit(r, Interval{tmin, tmax}, &hi)) {
                    id_out[i] = hi.object;
                    t_out[i] = hi.t;
                } else {
                    id_out[i] = -1;
                    t_out[i] = 0;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return 0;
}

// Full HitInfo of World.Hit for one ray (KATs).  out = point[3], normal[3], t, u, v, front.
int orc_hit_info(const rt_scene_desc *desc, const float *origin, const float *dir, float tmin,
                 float tmax, int32_t *id_out, float *out10) {
    Scene sc;
    if (!load_scene(desc, &sc)) return -1;
    HitInfo hi{};
    Ray r{ld3(origin), ld3(dir)};
    if (!world_hit(sc, r, Interval{tmin, tmax}, &hi)) {
        *id_out = -1;
        return 0;
    }
    *id_out = hi.object;
    float v[10] = {hi.point.x, hi.point.y, hi.point.z, hi.normal.x, hi.normal.y, hi.normal.z,
                   hi.t, hi.u, hi.v, hi.front_face ? 1.0f : 0.0f};
    memcpy(out10, v, sizeof v);
    return 0;
}

int orc_aabb_hit(const float *bmin, const float *bmax, const float *origin, const float *dir,
                 float tmin, float tmax) {
    Aabb a{{bmin[0], bmax[0]}, {bmin[1], bmax[1]}, {bmin[2], bmax[2]}};
    return aabb_hit(a, Ray{ld3(origin), ld3(dir)}, Interval{tmin, tmax}) ? 1 : 0;
}

void orc_reflect(const float *v, const float *n, float *out) {
    V3 r = reflect(ld3(v), ld3(n));
    out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void orc_refract(const float *uv, const float *n, float eta, float *out) {
    V3 r = refract(ld3(uv), ld3(n), eta);
    out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
float orc_reflectance(float cos_theta, float eta) { return reflectance(cos_theta, eta); }

void orc_texture(const rt_scene_desc *desc, uint32_t tex, float u, float v, const float *p, float *out) {
    Scene sc;
    if (!load_scene(desc, &sc)) return;
    V3 c = texture_value(sc, tex, u, v, ld3(p));
    out[0] = c.x, out[1] = c.y, out[2] = c.z;
}

void orc_encode_pixel(const float *mean, uint8_t *rgb) { encode_pixel(ld3(mean), rgb); }

// One Material.Scatter call on a World.Hit result, with the stream of (seed, pixel, sample).
// out = did_scatter, origin[3], dir[3], attenuation[3]
int orc_scatter(const rt_scene_desc *desc, const float *origin, const float *dir, uint64_t seed,
                uint32_t pixel, uint32_t sample, float *out10) {
    Scene sc;
    if (!load_scene(desc, &sc)) return -1;
    HitInfo hi{};
    Ray r{ld3(origin), ld3(dir)};
    if (!world_hit(sc, r, Interval{T_MIN, std::numeric_limits<float>::infinity()}, &hi)) return 1;
    Rng rng(seed, pixel, sample);
    Scatter s{};
    bool ok = material_scatter(sc, sc.materials[sc.material_of(hi.object)], r, hi, rng, &s);
    float v[10] = {ok ? 1.0f : 0.0f, s.ray.origin.x, s.ray.origin.y, s.ray.origin.z, s.ray.dir.x,
                   s.ray.dir.y, s.ray.dir.z, s.attenuation.x, s.attenuation.y, s.attenuation.z};
    memcpy(out10, v, sizeof v);
    return 0;
}

// Reference pin: Material.Scatter at the World.Hit of the ray with the uniforms of `feed` (four per block, layout at
// struct Rng) instead of the Philox stream.  out10 as orc_scatter, out3 = Material.Emit at the hit.
int orc_scatter_fed(const rt_scene_desc *desc, const float *origin, const float *dir, const float *feed,
                    int64_t n_feed, float *out10, float *emit3) {
    Scene sc;
    if (!load_scene(desc, &sc)) return -1;
    HitInfo hi{};
    Ray r{ld3(origin), ld3(dir)};
    if (!world_hit(sc, r, Interval{T_MIN, std::numeric_limits<float>::infinity()}, &hi)) return 1;
    Rng rng(feed, n_feed);
    Scatter s{};
    const rt_material &m = sc.materials[sc.material_of(hi.object)];
    bool ok = material_scatter(sc, m, r, hi, rng, &s);
    V3 e = material_emit(sc, m, hi);
    float v[10] = {ok ? 1.0f : 0.0f, s.ray.origin.x, s.ray.origin.y, s.ray.origin.z, s.ray.dir.x,
                   s.ray.dir.y, s.ray.dir.z, s.attenuation.x, s.attenuation.y, s.attenuation.z};
    memcpy(out10, v, sizeof v);
    emit3[0] = e.x, emit3[1] = e.y, emit3[2] = e.z;
    return 0;
}

// Reference pin: Ray.GetColor (ray.go:32-54) for an explicit ray, uniforms from `feed`.
int orc_get_color_fed(const rt_scene_desc *desc, const float *origin, const float *dir, const float *background,
                      int max_depth, int order, const float *feed, int64_t n_feed, float *out3) {
    Scene sc;
    if (!load_scene(desc, &sc)) return -1;
    World w{&sc, nullptr};
    Rng rng(feed, n_feed);
    Ray r{ld3(origin), ld3(dir)};
    V3 c = order == ORDER_ITERATIVE ? get_color_iterative(w, r, ld3(background), max_depth, rng)
                                    : get_color_recursive(w, r, ld3(background), max_depth, rng);
    out3[0] = c.x, out3[1] = c.y, out3[2] = c.z;
    return 0;
}

// Reference pin: Camera.GetRay (camera.go:265-299) for pixel (i, j), uniforms from `feed`.
int orc_get_ray_fed(const rt_camera *cam, int i, int j, const float *feed, int64_t n_feed, float *origin3, float *dir3) {
    if (!cam) return -1;
    Rng rng(feed, n_feed);
    Ray r = get_ray(*cam, rng, i, j);
    origin3[0] = r.origin.x, origin3[1] = r.origin.y, origin3[2] = r.origin.z;
    dir3[0] = r.dir.x, dir3[1] = r.dir.y, dir3[2] = r.dir.z;
    return 0;
}

// Camera.GetRay for pixels [pixel_begin, +n_pixels) x samples [sample_offset, +sample_count),
// sample-minor, same layout as rt_primary_rays.
int orc_primary_rays(const rt_camera *cam, uint64_t seed, int32_t sample_offset, int32_t sample_count,
                     int64_t pixel_begin, int64_t n_pixels, float *origins_out, float *dirs_out) {
    if (!cam) return -1;
    for (int64_t p = 0; p < n_pixels; p++) {
        int64_t pix = pixel_begin + p;
        int i = (int)(pix % cam->width), j = (int)(pix / cam->width);
        for (int k = 0; k < sample_count; k++) {
            Rng rng(seed, (uint32_t)pix, (uint32_t)(sample_offset + k));
            Ray r = get_ray(*cam, rng, i, j);
            float *o = origins_out + 3 * (p * sample_count + k), *d = dirs_out + 3 * (p * sample_count + k);
            o[0] = r.origin.x, o[1] = r.origin.y, o[2] = r.origin.z;
            d[0] = r.dir.x, d[1] = r.dir.y, d[2] = r.dir.z;
        }
    }
    return 0;
}

// Camera.Render's compute half (camera.go:198-222 + 254-263) on `threads` host threads over
// scanlines.  accum_out (nullable): W*H*3 float32 sums in sample order; rgb_out (nullable):
// resolved with 1/total_spp (total_spp = 0 -> sample_count).  Rows [row_begin,row_end) only
// (row_end = 0 -> height); other rows are left untouched.
int orc_render(const rt_scene_desc *desc, const rt_camera *cam, uint64_t seed, int32_t sample_offset,
               int32_t sample_count, int32_t total_spp, int mode, int order, uint64_t bvh_seed,
               int threads, int32_t row_begin, int32_t row_end, uint8_t *rgb_out, float *accum_out,
               orc_stats *stats) {
    Scene sc;
    if (!cam || !load_scene(desc, &sc)) return -1;
    RefBvh bvh;
    if (mode == MODE_REF_BVH && !sc.prims.empty()) {
        std::vector<int32_t> prims(sc.prims.size());
        for (size_t i = 0; i < prims.size(); i++) prims[i] = (int32_t)i;
        SplitMix rng{bvh_seed};
        bvh.root = ref_bvh_build(sc, bvh, prims, rng);
    }
    World w{&sc, mode == MODE_REF_BVH ? &bvh : nullptr};
    if (sample_count <= 0) sample_count = cam->spp;
    if (total_spp <= 0) total_spp = sample_count;
    if (row_end <= 0 || row_end > cam->height) row_end = cam->height;
    if (threads < 1) threads = 1;
    const int W = cam->width;
    V3 bg = ld3(cam->background);
    std::atomic<int> next_row{row_begin};
    std::vector<Counters> totals(threads);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        g_cnt = Counters{};
        for (;;) {
            int j = next_row.fetch_add(1);
            if (j >= row_end) break;
            for (int i = 0; i < W; i++) {
                // GetPixelColor, camera.go:254-263
                V3 sum = v3(0, 0, 0);
                uint32_t pix = (uint32_t)(j * (int64_t)W + i);
                for (int k = 0; k < sample_count; k++) {
                    Rng rng(seed, pix, (uint32_t)(sample_offset + k));
                    Ray r = get_ray(*cam, rng, i, j);
                    V3 s = order == ORDER_ITERATIVE
                               ? get_color_iterative(w, r, bg, cam->max_depth, rng)
                               : get_color_recursive(w, r, bg, cam->max_depth, rng);
                    sum = add(sum, s);
                    g_cnt.samples++;
                }
                if (accum_out) {
                    float *a = accum_out + 3 * (size_t)pix;
                    a[0] = sum.x, a[1] = sum.y, a[2] = sum.z;
                }
                if (rgb_out) {
                    V3 mean = scale(sum, 1.0f / (float)total_spp); // camera.go:261
                    encode_pixel(mean, rgb_out + 3 * (size_t)pix);
                }
            }
        }
        totals[tid] = g_cnt;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto &t : pool) t.join();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (auto &c : totals) {
            stats->samples += c.samples, stats->rays += c.rays, stats->box_tests += c.box_tests;
            stats->sphere_tests += c.sphere_tests, stats->hits += c.hits;
        }
        stats->seconds = secs;
        stats->threads = threads;
    }
    return 0;
}

// Resolve only (camera.go:261 + 212-214) — used to check rt_resolve_device.
void orc_resolve(const float *accum, int64_t n_pixels, int32_t total_spp, uint8_t *rgb_out) {
    for (int64_t p = 0; p < n_pixels; p++) {
        V3 mean = scale(ld3(accum + 3 * p), 1.0f / (float)total_spp);
        encode_pixel(mean, rgb_out + 3 * p);
    }
}

int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

} // extern "C"
