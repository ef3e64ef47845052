// hostsim.cpp — TEST-ONLY host compile of the per-ray device headers (rt_trace.h, rt_shade.h,
// rt_rng.h) and of the host BVH builder, so that traversal / shading / flattening logic can be
// checked against the oracle in this GPU-less container before GPU minutes are spent.
// It is NOT a CPU path of the product: librt_b200.so never contains or loads this code, and only
// `-m "not gpu"` tests build and call it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "../../include/rt_b200.h"
#include "../../raytracer_go_b200/csrc/bvh_build.h"
#include "../../raytracer_go_b200/csrc/rt_shade.h"

namespace {
struct HostScene {
    ScenePrims prims;
    FlatBvh bvh;
    std::vector<F4> mats;
    std::vector<DevImage> images;
    std::vector<std::vector<uint16_t>> texels;
    std::vector<DevPerlin> perlins;
    DevTex tex() const { return DevTex{images.data(), perlins.data()}; }
};

void load(const rt_scene_desc *d, int max_leaf, float origin_radius, HostScene *s) {
    double m[3], ext;
    load_scene_prims(d, &s->prims);
    double surf;
    compute_scene_center(s->prims, m, &ext, &surf);
    // derived radius as rt_scene_create + a camera / ray batch anywhere within the scene's extent
    float R = origin_radius > 0 ? origin_radius : (d->ray_origin_radius > 0 ? d->ray_origin_radius : (float)(2 * ext + surf));
    build_flat_bvh(s->prims, R, max_leaf, &s->bvh, nullptr, true); // with the leaf-start chains, as the kernels get them
    pack_materials(d, &s->mats);
    s->images.resize(d->n_images);
    s->texels.resize(d->n_images);
    for (uint32_t i = 0; i < d->n_images; i++) {
        const rt_image &im = d->images[i];
        size_t n = (size_t)(im.w > 0 ? im.w : 0) * (size_t)(im.h > 0 ? im.h : 0);
        s->texels[i].resize(n * 4);
        for (size_t k = 0; k < n; k++)
            for (int c = 0; c < 3; c++) s->texels[i][4 * k + c] = im.rgb16[3 * k + c];
        s->images[i].texels = s->texels[i].data(), s->images[i].w = im.w, s->images[i].h = im.h;
    }
    s->perlins.resize(d->n_perlins);
    for (uint32_t i = 0; i < d->n_perlins; i++) {
        for (int k = 0; k < 256; k++) {
            s->perlins[i].vec[k].x = d->perlins[i].vec[k][0], s->perlins[i].vec[k].y = d->perlins[i].vec[k][1];
            s->perlins[i].vec[k].z = d->perlins[i].vec[k][2], s->perlins[i].vec[k].w = 0;
        }
        memcpy(s->perlins[i].perm_x, d->perlins[i].perm_x, 256), memcpy(s->perlins[i].perm_y, d->perlins[i].perm_y, 256);
        memcpy(s->perlins[i].perm_z, d->perlins[i].perm_z, 256);
    }
}
// Emit + Scatter at a HitRec (sphere or quad slot), like the kernel does.
bool shade_at(const HostScene &s, const HitRec &h, PathRng &rng, V3 &o, V3 &dir, V3 &atten, V3 &emitted) {
    if (h.slot & RT_HIT_QUAD) {
        const F4 *q = s.bvh.quad.data() + (size_t)RT_QUAD_F4 * (h.slot & ~RT_HIT_QUAD);
        const uint32_t mi = as_uint(q[1].w);
        return shade_hit_quad(s.mats[2 * mi], s.mats[2 * mi + 1], s.tex(), q, h.t, rng, o, dir, atten, emitted);
    }
    const int mi = s.bvh.meta[h.slot].y;
    return shade_hit(s.mats[2 * mi], s.mats[2 * mi + 1], s.tex(), s.bvh.sph[h.slot], h.t, rng, o, dir, atten, emitted);
}
} // namespace

extern "C" {

int hs_bvh_stats(const rt_scene_desc *d, int max_leaf, float origin_radius, uint64_t *n_nodes, uint64_t *n_slots,
                 uint32_t *max_depth, float *pad_min, float *pad_max) {
    HostScene s;
    load(d, max_leaf, origin_radius, &s);
    *n_nodes = s.bvh.nodes.size() / 2, *n_slots = s.bvh.sph.size(), *max_depth = s.bvh.max_depth;
    *pad_min = s.bvh.pad_min, *pad_max = s.bvh.pad_max;
    return 0;
}

// FNV-1a over every array of the flattened BVH: the parallel build must give the same bytes for any
// thread count (RT_B200_BVH_THREADS), and refit must not depend on it either.
static uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}
uint64_t hs_bvh_hash(const rt_scene_desc *d, int max_leaf, float origin_radius, float refit_radius) {
    HostScene s;
    load(d, max_leaf, origin_radius, &s);
    if (refit_radius > 0) refit_flat_bvh(s.prims, refit_radius, &s.bvh);
    uint64_t h = 1469598103934665603ull;
    const FlatBvh &b = s.bvh;
    h = fnv(h, b.nodes.data(), b.nodes.size() * sizeof(F4));
    h = fnv(h, b.dev_nodes.data(), b.nodes.size() * sizeof(F4)); // the tree's own nodes (the leaf-start walk pairs follow them)
    h = fnv(h, b.sph.data(), b.sph.size() * sizeof(F4));
    h = fnv(h, b.meta.data(), b.meta.size() * sizeof(I2));
    h = fnv(h, b.sph_prim.data(), b.sph_prim.size() * 4);
    h = fnv(h, b.quad.data(), b.quad.size() * sizeof(F4));
    h = fnv(h, b.quad_prim.data(), b.quad_prim.size() * 4);
    h = fnv(h, &b.root_ref, 4), h = fnv(h, &b.max_depth, 4), h = fnv(h, &b.pad_min, 4), h = fnv(h, &b.pad_max, 4);
    return h;
}

int hs_trace(const rt_scene_desc *d, int max_leaf, float origin_radius, const float *origins, const float *dirs,
             int64_t n, float tmin, float tmax, int32_t *id_out, float *t_out, uint64_t *box_tests,
             uint64_t *sphere_tests) {
    HostScene s;
    load(d, max_leaf, origin_radius, &s);
    WorkCounters wc{0, 0};
    for (int64_t i = 0; i < n; i++) {
        LocalStack<64> stack;
        HitRec h;
        trace_closest<LocalStack<64>, true, true>(s.bvh.dev_nodes.data(), s.bvh.sph.data(), s.bvh.meta.data(), s.bvh.root_ref,
                                                  v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]),
                                                  v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), tmin, tmax, stack, h, &wc,
                                                  s.bvh.quad.data());
        if (h.slot == RT_REF_NONE) id_out[i] = -1, t_out[i] = 0;
        else id_out[i] = slot_object_id(h.slot, s.bvh.meta.data(), s.bvh.quad.data()), t_out[i] = h.t;
    }
    if (box_tests) *box_tests = wc.box_tests;
    if (sphere_tests) *sphere_tests = wc.sphere_tests;
    return 0;
}

// hs_trace with leaf start: ray i starts its traversal at the leaf of sphere slot start_slots[i] % n_slots (rt_trace.h).
// A leaf's chain plus the leaf itself cover the whole tree, so ANY start slot must give the answer of the root start;
// the work counters show what the right start saves.
int hs_trace_leaf_start(const rt_scene_desc *d, int max_leaf, float origin_radius, const float *origins, const float *dirs,
                        const uint32_t *start_slots, int64_t n, float tmin, float tmax, int32_t *id_out, float *t_out,
                        uint64_t *box_tests, uint64_t *sphere_tests) {
    HostScene s;
    load(d, max_leaf, origin_radius, &s);
    WorkCounters wc{0, 0};
    const size_t n_slots = s.bvh.sph.size(), n_q = s.bvh.quad_prim.size();
    for (int64_t i = 0; i < n; i++) {
        LocalStack<64> stack;
        HitRec h;
        uint32_t start = RT_REF_NONE;
        if (n_slots + n_q) {
            const size_t k = start_slots[i] % (n_slots + n_q);
            start = k < n_slots ? s.bvh.sph_chain[k] : s.bvh.quad_chain[k - n_slots];
        }
        trace_closest<LocalStack<64>, true, true>(s.bvh.dev_nodes.data(), s.bvh.sph.data(), s.bvh.meta.data(), s.bvh.root_ref,
                                                  v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]),
                                                  v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), tmin, tmax, stack, h, &wc,
                                                  s.bvh.quad.data(), s.bvh.chains.data(), start);
        if (h.slot == RT_REF_NONE) id_out[i] = -1, t_out[i] = 0;
        else id_out[i] = slot_object_id(h.slot, s.bvh.meta.data(), s.bvh.quad.data()), t_out[i] = h.t;
    }
    if (box_tests) *box_tests = wc.box_tests;
    if (sphere_tests) *sphere_tests = wc.sphere_tests;
    return 0;
}

// The megakernel's per-path loop + reduce + resolve, serially, in the same operation order.
int hs_render(const rt_scene_desc *d, const rt_camera *cam, uint64_t seed, int32_t sample_offset, int32_t sample_count,
              int32_t total_spp, int max_leaf, uint8_t *rgb_out, float *accum_out) {
    HostScene s;
    load(d, max_leaf, 0, &s);
    DevCamera c = make_dev_camera(*cam);
    const int64_t n_pix = (int64_t)cam->width * cam->height;
    for (int64_t pix = 0; pix < n_pix; pix++) {
        V3 sum = v3(0, 0, 0);
        for (int k = 0; k < sample_count; k++) {
            PathRng rng;
            rng.init(seed, (uint32_t)pix, (uint32_t)(sample_offset + k));
            V3 o, dir;
            generate_ray(c, rng, (int)(pix % cam->width), (int)(pix / cam->width), o, dir);
            V3 thr = v3(1, 1, 1), rad = v3(0, 0, 0);
            uint32_t start = RT_REF_NONE; // leaf start, as render_kernel: secondary rays start at the leaf they leave
            for (int depth = 0; depth < c.max_depth;) {
                LocalStack<64> stack;
                HitRec h;
                trace_closest<LocalStack<64>, false, true>(s.bvh.dev_nodes.data(), s.bvh.sph.data(), s.bvh.meta.data(),
                                                           s.bvh.root_ref, o, dir, 0.001f, INFINITY, stack, h, nullptr,
                                                           s.bvh.quad.data(), s.bvh.chains.data(), start);
                if (h.slot == RT_REF_NONE) {
                    rad = rad + thr * c.background;
                    break;
                }
                start = (h.slot & RT_HIT_QUAD) ? s.bvh.quad_chain[h.slot & ~RT_HIT_QUAD] : s.bvh.sph_chain[h.slot];
                V3 atten, emitted;
                bool sc = shade_at(s, h, rng, o, dir, atten, emitted);
                rad = rad + thr * emitted;
                if (!sc) break;
                thr = thr * atten;
                depth++;
            }
            sum = sum + rad;
        }
        if (accum_out) accum_out[3 * pix] = sum.x, accum_out[3 * pix + 1] = sum.y, accum_out[3 * pix + 2] = sum.z;
        if (rgb_out) resolve_pixel(sum, 1.0f / (float)total_spp, rgb_out + 3 * pix);
    }
    return 0;
}

int hs_primary_rays(const rt_camera *cam, uint64_t seed, int32_t sample_offset, int32_t sample_count, int64_t pixel_begin,
                    int64_t n_pixels, float *origins, float *dirs) {
    DevCamera c = make_dev_camera(*cam);
    for (int64_t p = 0; p < n_pixels; p++)
        for (int k = 0; k < sample_count; k++) {
            int64_t pix = pixel_begin + p;
            PathRng rng;
            rng.init(seed, (uint32_t)pix, (uint32_t)(sample_offset + k));
            V3 o, dir;
            generate_ray(c, rng, (int)(pix % cam->width), (int)(pix / cam->width), o, dir);
            float *po = origins + 3 * (p * sample_count + k), *pd = dirs + 3 * (p * sample_count + k);
            po[0] = o.x, po[1] = o.y, po[2] = o.z, pd[0] = dir.x, pd[1] = dir.y, pd[2] = dir.z;
        }
    return 0;
}
}

// Per-segment traversal work of real paths (design exploration: SIMT trip-count statistics).
// For path p = pixel-major/sample-minor index, writes up to max_seg records (inner iterations,
// sphere tests, depth) and returns the number written.
extern "C" int64_t hs_segment_stats(const rt_scene_desc *d, const rt_camera *cam, uint64_t seed, int32_t spp,
                                    int64_t pixel_begin, int64_t n_pixels, int max_leaf, int32_t *iters,
                                    int32_t *sph_tests, int32_t *depths, int64_t max_seg) {
    HostScene s;
    load(d, max_leaf, 0, &s);
    DevCamera c = make_dev_camera(*cam);
    int64_t n = 0;
    for (int64_t pp = 0; pp < n_pixels; pp++) {
        int64_t pix = pixel_begin + pp;
        for (int k = 0; k < spp; k++) {
            PathRng rng;
            rng.init(seed, (uint32_t)pix, (uint32_t)k);
            V3 o, dir;
            generate_ray(c, rng, (int)(pix % cam->width), (int)(pix / cam->width), o, dir);
            for (int depth = 0; depth < c.max_depth;) {
                LocalStack<64> stack;
                HitRec h;
                WorkCounters wc{0, 0};
                trace_closest<LocalStack<64>, true>(s.bvh.dev_nodes.data(), s.bvh.sph.data(), s.bvh.meta.data(),
                                                    s.bvh.root_ref, o, dir, 0.001f, INFINITY, stack, h, &wc);
                if (n < max_seg) iters[n] = (int32_t)(wc.box_tests / 2), sph_tests[n] = (int32_t)wc.sphere_tests, depths[n] = depth, n++;
                if (h.slot == RT_REF_NONE) break;
                const int mi = s.bvh.meta[h.slot].y;
                V3 atten, emitted;
                if (!shade_hit(s.mats[2 * mi], s.mats[2 * mi + 1], s.tex(), s.bvh.sph[h.slot], h.t, rng, o, dir, atten, emitted)) break;
                depth++;
            }
        }
    }
    return n;
}

// Event trace of the traversal of real path segments (design exploration only): per ray a
// sequence of tokens, +n = n consecutive inner-node steps, -n = a leaf with n sphere tests, 0 = end.
extern "C" int64_t hs_traversal_events(const rt_scene_desc *d, const rt_camera *cam, uint64_t seed, int32_t spp,
                                       int64_t pixel_begin, int64_t n_pixels, int max_leaf, int8_t *tokens,
                                       int64_t max_tokens, int64_t *n_rays_out) {
    HostScene s;
    load(d, max_leaf, 0, &s);
    DevCamera c = make_dev_camera(*cam);
    int64_t n = 0, rays = 0;
    const F4 *nodes = s.bvh.dev_nodes.data();
    for (int64_t pp = 0; pp < n_pixels; pp++) {
        int64_t pix = pixel_begin + pp;
        for (int k = 0; k < spp; k++) {
            PathRng rng;
            rng.init(seed, (uint32_t)pix, (uint32_t)k);
            V3 o, dir;
            generate_ray(c, rng, (int)(pix % cam->width), (int)(pix / cam->width), o, dir);
            for (int depth = 0; depth < c.max_depth;) {
                // same loop as trace_closest, with event logging
                const V3 inv = v3(cull_rcp(dir.x), cull_rcp(dir.y), cull_rcp(dir.z));
                const V3 noi = v3(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
                const V3 ainv = v3(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
                const float a = lensq(dir);
                float tbest = INFINITY;
                uint32_t best = RT_REF_NONE, ref = s.bvh.root_ref;
                LocalStack<64> stack;
                stack.reset();
                if (n + 130 > max_tokens) goto out;
                for (;;) {
                    int run = 0;
                    while (!(ref & RT_LEAF)) {
                        const F4 l0 = nodes[2 * ref], l1 = nodes[2 * ref + 1], r0 = nodes[2 * ref + 2], r1 = nodes[2 * ref + 3];
                        float tl, tr;
                        const bool hl = box_test(l0, l1, inv, noi, ainv, 0.001f, tbest, tl), hr = box_test(r0, r1, inv, noi, ainv, 0.001f, tbest, tr);
                        const uint32_t lref = as_uint(l0.w), rref = as_uint(r0.w);
                        if (hl && hr) {
                            const bool lf = tl <= tr;
                            stack.push(lf ? rref : lref);
                            ref = lf ? lref : rref;
                        } else if (hl) ref = lref;
                        else if (hr) ref = rref;
                        else ref = stack.pop();
                        if (++run == 127) tokens[n++] = 127, run = 0;
                    }
                    if (run) tokens[n++] = (int8_t)run;
                    if (ref == RT_REF_NONE) break;
                    const uint32_t first = (ref & ~RT_LEAF) >> 3, count = (ref & 7u) + 1;
                    for (uint32_t q = first; q < first + count; q++) {
                        float t;
                        if (!sphere_candidate(s.bvh.sph[q], o, dir, a, 0.001f, t)) continue;
                        if (t < tbest) tbest = t, best = q;
                        else if (t == tbest && best != RT_REF_NONE && s.bvh.meta[q].x < s.bvh.meta[best].x) best = q;
                    }
                    tokens[n++] = (int8_t)(-(int)count);
                    ref = stack.pop();
                    if (n + 4 > max_tokens) goto out;
                }
                tokens[n++] = 0;
                rays++;
                if (best == RT_REF_NONE) break;
                const int mi = s.bvh.meta[best].y;
                V3 atten, emitted;
                if (!shade_hit(s.mats[2 * mi], s.mats[2 * mi + 1], s.tex(), s.bvh.sph[best], tbest, rng, o, dir, atten, emitted)) break;
                depth++;
            }
        }
    }
out:
    *n_rays_out = rays;
    return n;
}

// As hs_traversal_events, but also records each ray's origin and direction (6 floats per ray) so
// that ray-sorting policies can be evaluated offline.
extern "C" int64_t hs_traversal_events_rays(const rt_scene_desc *d, const rt_camera *cam, uint64_t seed, int32_t spp,
                                            int64_t pixel_begin, int64_t n_pixels, int max_leaf, int8_t *tokens,
                                            int64_t max_tokens, float *rays_out, int64_t max_rays, int64_t *n_rays_out) {
    HostScene s;
    load(d, max_leaf, 0, &s);
    DevCamera c = make_dev_camera(*cam);
    int64_t n = 0, rays = 0;
    const F4 *nodes = s.bvh.dev_nodes.data();
    for (int64_t pp = 0; pp < n_pixels; pp++) {
        int64_t pix = pixel_begin + pp;
        for (int k = 0; k < spp; k++) {
            PathRng rng;
            rng.init(seed, (uint32_t)pix, (uint32_t)k);
            V3 o, dir;
            generate_ray(c, rng, (int)(pix % cam->width), (int)(pix / cam->width), o, dir);
            for (int depth = 0; depth < c.max_depth;) {
                if (n + 130 > max_tokens || rays >= max_rays) goto out;
                float *ro = rays_out + 7 * rays;
                ro[0] = o.x, ro[1] = o.y, ro[2] = o.z, ro[3] = dir.x, ro[4] = dir.y, ro[5] = dir.z, ro[6] = (float)depth;
                const V3 inv = v3(cull_rcp(dir.x), cull_rcp(dir.y), cull_rcp(dir.z));
                const V3 noi = v3(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
                const V3 ainv = v3(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
                const float a = lensq(dir);
                float tbest = INFINITY;
                uint32_t best = RT_REF_NONE, ref = s.bvh.root_ref;
                LocalStack<64> stack;
                stack.reset();
                for (;;) {
                    int run = 0;
                    while (!(ref & RT_LEAF)) {
                        const F4 l0 = nodes[2 * ref], l1 = nodes[2 * ref + 1], r0 = nodes[2 * ref + 2], r1 = nodes[2 * ref + 3];
                        float tl, tr;
                        const bool hl = box_test(l0, l1, inv, noi, ainv, 0.001f, tbest, tl), hr = box_test(r0, r1, inv, noi, ainv, 0.001f, tbest, tr);
                        const uint32_t lref = as_uint(l0.w), rref = as_uint(r0.w);
                        if (hl && hr) {
                            const bool lf = tl <= tr;
                            stack.push(lf ? rref : lref);
                            ref = lf ? lref : rref;
                        } else if (hl) ref = lref;
                        else if (hr) ref = rref;
                        else ref = stack.pop();
                        if (++run == 127) tokens[n++] = 127, run = 0;
                    }
                    if (run) tokens[n++] = (int8_t)run;
                    if (ref == RT_REF_NONE) break;
                    const uint32_t first = (ref & RT_LEAF_SLOT_MASK) >> 3, count = (ref & 7u) + 1;
                    for (uint32_t q = first; q < first + count; q++) {
                        float t;
                        if (!sphere_candidate(s.bvh.sph[q], o, dir, a, 0.001f, t)) continue;
                        if (t < tbest) tbest = t, best = q;
                    }
                    tokens[n++] = (int8_t)(-(int)count);
                    ref = stack.pop();
                }
                tokens[n++] = 0;
                rays++;
                if (best == RT_REF_NONE) break;
                const int mi = s.bvh.meta[best].y;
                V3 atten, emitted;
                if (!shade_hit(s.mats[2 * mi], s.mats[2 * mi + 1], s.tex(), s.bvh.sph[best], tbest, rng, o, dir, atten, emitted)) break;
                depth++;
            }
        }
    }
out:
    *n_rays_out = rays;
    return n;
}

// Design exploration ("per-pixel beam"): for warps of 32 consecutive samples of one
// pixel, walk the flat BVH once with the interval slab test that is exact for {origin box} x {direction box}
// and report, per warp: node pairs visited, spheres in the visited leaves (the candidate list), how many of
// them the rays would test if the list were sorted by beam entry distance and abandoned once every ray's
// closest hit lies in front of the next candidate, and (for comparison) the per-ray traversal's node pairs
// and sphere tests averaged over the 32 rays.  out = 5 floats per warp.
extern "C" int64_t hs_beam_stats(const rt_scene_desc *d, const rt_camera *cam, uint64_t seed, int64_t pixel_begin,
                                 int64_t n_pixels, int64_t pixel_stride, int max_leaf, float *out) {
    HostScene s;
    load(d, max_leaf, 0, &s);
    DevCamera c = make_dev_camera(*cam);
    const F4 *nodes = s.bvh.nodes.data(); // (min.xyz, ref)(max.xyz, 0): the builder's padded boxes
    int64_t n = 0;
    for (int64_t pp = 0; pp < n_pixels; pp++) {
        const int64_t pix = pixel_begin + pp * pixel_stride;
        V3 o[32], dir[32];
        float omin[3] = {INFINITY, INFINITY, INFINITY}, omax[3] = {-INFINITY, -INFINITY, -INFINITY};
        float dmin[3] = {INFINITY, INFINITY, INFINITY}, dmax[3] = {-INFINITY, -INFINITY, -INFINITY};
        float tbest[32];
        double ray_pairs = 0, ray_sph = 0;
        for (int k = 0; k < 32; k++) {
            PathRng rng;
            rng.init(seed, (uint32_t)pix, (uint32_t)k);
            generate_ray(c, rng, (int)(pix % cam->width), (int)(pix / cam->width), o[k], dir[k]);
            const float ov[3] = {o[k].x, o[k].y, o[k].z}, dv[3] = {dir[k].x, dir[k].y, dir[k].z};
            for (int a = 0; a < 3; a++) {
                omin[a] = std::min(omin[a], ov[a]), omax[a] = std::max(omax[a], ov[a]);
                dmin[a] = std::min(dmin[a], dv[a]), dmax[a] = std::max(dmax[a], dv[a]);
            }
            LocalStack<64> stack;
            HitRec h;
            WorkCounters wc{0, 0};
            trace_closest<LocalStack<64>, true>(s.bvh.dev_nodes.data(), s.bvh.sph.data(), s.bvh.meta.data(), s.bvh.root_ref,
                                                o[k], dir[k], 0.001f, INFINITY, stack, h, &wc);
            tbest[k] = h.slot == RT_REF_NONE ? INFINITY : h.t;
            ray_pairs += (double)wc.box_tests / 2, ray_sph += (double)wc.sphere_tests;
        }
        // entry distance of the beam into a box, or a negative value when it cannot touch it
        auto beam_enter = [&](const F4 &lo, const F4 &hi) -> float {
            const float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
            float t0 = 0.001f, t1 = INFINITY;
            for (int a = 0; a < 3; a++) {
                const float A = h[a] - omin[a], B = l[a] - omax[a]; // omin + t dmin <= hi ; omax + t dmax >= lo
                if (dmin[a] > 0) t1 = std::min(t1, A / dmin[a]);
                else if (dmin[a] < 0) t0 = std::max(t0, A / dmin[a]);
                else if (A < 0) return -1.0f;
                if (dmax[a] > 0) t0 = std::max(t0, B / dmax[a]);
                else if (dmax[a] < 0) t1 = std::min(t1, B / dmax[a]);
                else if (B > 0) return -1.0f;
            }
            return t0 <= t1 ? t0 : -1.0f;
        };
        std::vector<uint32_t> stack;
        std::vector<std::pair<float, uint32_t>> cand; // (beam entry distance of the leaf, sphere slot)
        int pairs = 0;
        auto visit_ref = [&](uint32_t ref, float t_enter) {
            if (ref & RT_LEAF) {
                if (ref == RT_REF_NONE || (ref & RT_LEAF_QUAD)) return;
                const uint32_t first = (ref & RT_LEAF_SLOT_MASK) >> 3, count = (ref & 7u) + 1;
                for (uint32_t q = first; q < first + count; q++) cand.push_back({t_enter, q});
            } else {
                stack.push_back(ref);
            }
        };
        visit_ref(s.bvh.root_ref, 0.0f);
        while (!stack.empty()) {
            const uint32_t ref = stack.back();
            stack.pop_back();
            pairs++;
            for (int ch = 0; ch < 2; ch++) {
                const F4 &lo = nodes[2 * (ref + ch)], &hi = nodes[2 * (ref + ch) + 1];
                const float t = beam_enter(lo, hi);
                if (t >= 0) visit_ref(as_uint(lo.w), t);
            }
        }
        std::sort(cand.begin(), cand.end());
        float worst = 0; // the farthest closest-hit among the 32 rays (infinite if one of them misses)
        for (int k = 0; k < 32; k++) worst = std::max(worst, tbest[k]);
        int tested = 0;
        for (auto &cd : cand) {
            if (cd.first > worst) break;
            tested++;
        }
        out[5 * n] = (float)pairs, out[5 * n + 1] = (float)cand.size(), out[5 * n + 2] = (float)tested;
        out[5 * n + 3] = (float)(ray_pairs / 32), out[5 * n + 4] = (float)(ray_sph / 32);
        n++;
    }
    return n;
}

// The per-pixel candidate lists of the primary stage (rt_kernels.cuh: pixel_candidates_kernel) replayed on the host
// with the very functions the kernel calls (pixel_beam, beam_candidates): for `n_pixels` pixels starting at
// pixel_begin (stride pixel_stride) and `spp` camera rays each, checks that (1) the closest hit of every ray is in the
// pixel's list and (2) trace_candidates over the list returns exactly what the tree traversal returns.
// out[0] = pixels, out[1] = sum of list lengths, out[2] = longest list, out[3] = rays whose hit is missing from the
// list, out[4] = rays where the two traces differ, out[5] = lists longer than `cap`.
extern "C" int hs_pixel_candidates(const rt_scene_desc *d, const rt_camera *cam, uint64_t seed, int64_t pixel_begin,
                                   int64_t n_pixels, int64_t pixel_stride, int32_t spp, int max_leaf, uint32_t cap, double *out) {
    HostScene s;
    load(d, max_leaf, 0, &s);
    DevCamera c = make_dev_camera(*cam);
    const bool quads = !s.bvh.quad_prim.empty();
    for (int k = 0; k < 6; k++) out[k] = 0;
    std::vector<uint32_t> list(4096);
    for (int64_t pp = 0; pp < n_pixels; pp++) {
        const int64_t pix = pixel_begin + pp * pixel_stride;
        const int i = (int)(pix % cam->width), j = (int)(pix / cam->width);
        const Beam b = pixel_beam(c, i, j);
        const uint32_t n = quads ? beam_candidates<true>(s.bvh.dev_nodes.data(), s.bvh.root_ref, b, list.data(), (uint32_t)list.size())
                                 : beam_candidates<false>(s.bvh.dev_nodes.data(), s.bvh.root_ref, b, list.data(), (uint32_t)list.size());
        out[0] += 1, out[1] += n, out[2] = std::max(out[2], (double)n);
        if (n > cap) out[5] += 1;
        if (n > list.size()) return -1;
        for (int k = 0; k < spp; k++) {
            PathRng rng;
            rng.init(seed, (uint32_t)pix, (uint32_t)k);
            V3 o, dir;
            generate_ray(c, rng, i, j, o, dir);
            LocalStack<64> stack;
            HitRec h, hc;
            WorkCounters wc{0, 0};
            if (quads) {
                trace_closest<LocalStack<64>, false, true>(s.bvh.dev_nodes.data(), s.bvh.sph.data(), s.bvh.meta.data(), s.bvh.root_ref, o, dir,
                                                           0.001f, INFINITY, stack, h, &wc, s.bvh.quad.data());
                trace_candidates<false, true>(list.data(), n, s.bvh.sph.data(), s.bvh.meta.data(), s.bvh.quad.data(), o, dir, 0.001f, INFINITY, hc, &wc);
            } else {
                trace_closest<LocalStack<64>, false, false>(s.bvh.dev_nodes.data(), s.bvh.sph.data(), s.bvh.meta.data(), s.bvh.root_ref, o, dir,
                                                            0.001f, INFINITY, stack, h, &wc);
                trace_candidates<false, false>(list.data(), n, s.bvh.sph.data(), s.bvh.meta.data(), nullptr, o, dir, 0.001f, INFINITY, hc, &wc);
            }
            if (h.slot != RT_REF_NONE && std::find(list.begin(), list.begin() + n, h.slot) == list.begin() + n) out[3] += 1;
            if (h.slot != hc.slot || (h.slot != RT_REF_NONE && as_uint(h.t) != as_uint(hc.t))) out[4] += 1;
        }
    }
    return 0;
}

// Property behind the candidate lists, at the level of ONE box: whenever the kernels' slab test (box_test) accepts a
// ray drawn from a beam, the beam test (beam_box_test) accepts the box.  Random boxes (incl. flat and huge ones), random
// beams (origin box x direction box, narrow like a pixel's or wide, some straddling zero on an axis), `rays` rays per
// beam.  Returns the number of violations; *accepted = (box, beam) pairs in which at least one ray was accepted.
extern "C" int64_t hs_beam_box_property(uint64_t seed, int64_t n_pairs, int rays, int64_t *accepted) {
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { // xorshift64*, uniform in [0, 1)
        st ^= st >> 12, st ^= st << 25, st ^= st >> 27;
        return (float)((st * 0x2545F4914F6CDD1Dull) >> 40) * (1.0f / 16777216.0f);
    };
    auto sym = [&](float s) { return (rnd() * 2 - 1) * s; };
    int64_t bad = 0, acc = 0;
    for (int64_t k = 0; k < n_pairs; k++) {
        const float scale = powf(10.0f, sym(2.0f));
        F4 c = {sym(20) * scale, sym(20) * scale, sym(20) * scale, 0}, h = {rnd() * 3 * scale, rnd() * 3 * scale, rnd() * 3 * scale, 0};
        if (k % 7 == 0) h.y = 0;                 // flat box
        if (k % 11 == 0) h.x = h.z = 1000 * scale; // the ground
        Beam b;
        const V3 o = v3(sym(30) * scale, sym(30) * scale, sym(30) * scale);
        const float ow = (k % 3 == 0) ? 0.0f : rnd() * 0.1f * scale;
        V3 d = v3(c.x - o.x + sym(2) * h.x, c.y - o.y + sym(2) * h.y, c.z - o.z + sym(2) * h.z); // aimed near the box
        if (k % 5 == 0) d = v3(sym(1), sym(1), sym(1));
        const float dw = (k % 4 == 0 ? 0.5f : 0.002f) * (fabsf(d.x) + fabsf(d.y) + fabsf(d.z)) * rnd();
        b.olo = v3(o.x - ow, o.y - ow, o.z - ow), b.ohi = v3(o.x + ow, o.y + ow, o.z + ow);
        b.dlo = v3(d.x - dw, d.y - dw, d.z - dw), b.dhi = v3(d.x + dw, d.y + dw, d.z + dw);
        const bool beam_hit = beam_box_test(c, h, b);
        bool any = false;
        for (int r = 0; r < rays; r++) {
            const V3 ro = v3(b.olo.x + rnd() * (b.ohi.x - b.olo.x), b.olo.y + rnd() * (b.ohi.y - b.olo.y), b.olo.z + rnd() * (b.ohi.z - b.olo.z));
            V3 rd = v3(b.dlo.x + rnd() * (b.dhi.x - b.dlo.x), b.dlo.y + rnd() * (b.dhi.y - b.dlo.y), b.dlo.z + rnd() * (b.dhi.z - b.dlo.z));
            if (r == 0) rd = b.dlo; // corners too
            if (r == 1) rd = b.dhi;
            const V3 inv = v3(cull_rcp(rd.x), cull_rcp(rd.y), cull_rcp(rd.z));
            const V3 noi = v3(-(ro.x * inv.x), -(ro.y * inv.y), -(ro.z * inv.z));
            const V3 ainv = v3(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
            float tn;
            if (box_test(c, h, inv, noi, ainv, 0.001f, INFINITY, tn)) any = true;
        }
        if (any) acc++;
        if (any && !beam_hit) bad++;
    }
    if (accepted) *accepted = acc;
    return bad;
}
