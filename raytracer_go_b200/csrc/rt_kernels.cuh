// rt_kernels.cuh — the sm_100a kernels of librt_b200.so and the device-side scene they read.
//
//   primary_stage_kernel  first segment of every path, traced as coherent warps (two-stage mode, default)
//   render_kernel       persistent megakernel: one lane = one (pixel, sample) path at a time
//   reduce_kernel       ordered FP32 accumulation of a pass's per-sample radiances (camera.go:255-260)
//   resolve_kernel      1/spp, sqrt, clamp, *255.999, truncate (camera.go:261, vec3.go:141-166)
//   add_kernel          accumulator += accumulator (multi-GPU gather)
//   trace_kernel        rt_trace parity hook (World.Hit as a query)
//   primary_kernel      rt_primary_rays parity hook (Camera.GetRay)
//
// Included once, by rt_b200.cu.  Compiled with -fmad=false (rt_math.h).
#ifndef RT_KERNELS_CUH
#define RT_KERNELS_CUH

#include <cuda_runtime.h>

#include <type_traits>

#include "rt_math.h"
#include "rt_rng.h"
#include "rt_shade.h"
#include "rt_trace.h"

// ---------------------------------------------------------------------------------------------
// device-side scene
// ---------------------------------------------------------------------------------------------
struct DevScene {
    const F4 *nodes;
    const F4 *sph;
    const I2 *meta;
    const F4 *mats;
    DevTex tex;      // images and Perlin tables (global memory)
    const F4 *quads; // RT_QUAD_F4 x F4 per quad slot
    // leaf start (rt_trace.h, bvh_build.h): chains == nullptr when the scene was built without it
    const uint32_t *chains, *sph_chain, *quad_chain;
    uint32_t n_chain_words;
    uint32_t root_ref, n_nodes, n_slots, n_mats, n_quad_slots; // n_nodes counts the walk pairs too
    uint32_t stack_depth; // entries per thread for the shared-memory stack
};

#define RT_SMEM_NODE_STRIDE 40 /* bytes per node in the shared-memory copy (rt_trace.h: NSTRIDE) */
// Shared-memory pairs transposed for FFMA2 (rt_trace.h: box_test_pair): 0 never, 1 always, 2 (default) only in the
// instantiations with quads.  Measured (profiles/r02h): the 9 FFMA2 of a pair cost what the 18 FFMA they replace do
// (a 64-bit operand read takes two issue cycles), so the sphere scenes gain nothing (C2 -1.5 %, C3 -0.8 %); the
// Cornell box, whose leaves run the longer quad test, gains 3.9 %.
#ifndef RT_PACKED_PAIRS
#define RT_PACKED_PAIRS 2
#endif
#define RT_PACKED(SMEM, QUADS) ((SMEM) && (RT_PACKED_PAIRS == 1 || (RT_PACKED_PAIRS == 2 && (QUADS))))
static size_t scene_smem_bytes(const DevScene &s) {
    return (size_t)s.n_nodes * RT_SMEM_NODE_STRIDE + (size_t)s.n_slots * 16 + (size_t)s.n_mats * 32 +
           (size_t)s.n_quad_slots * 16 * RT_QUAD_F4 + (size_t)s.n_slots * 8 +
           (s.chains ? ((size_t)s.n_chain_words + s.n_slots + s.n_quad_slots) * 4 : 0);
}

#define RT_LOCAL_STACK 64
#define RT_N_STATS 6 /* device counters: rays, hits, box tests, sphere tests, survivors, (spare) */

// Stage the scene arrays in shared memory (LDG.128 -> STS.128), return the carved pointers.
struct SmemScene {
    const F4 *nodes, *sph, *mats, *quads;
    const I2 *meta;
    const uint32_t *chains, *sph_chain, *quad_chain;
    uint32_t *stack;
};

// Chain of the leaf that holds hit slot `slot` (leaf start), RT_REF_NONE without chains.
__device__ __forceinline__ uint32_t chain_of_slot(const uint32_t *sph_chain, const uint32_t *quad_chain, uint32_t slot) {
    if (sph_chain == nullptr) return RT_REF_NONE;
    return (slot & RT_HIT_QUAD) ? quad_chain[slot & ~RT_HIT_QUAD] : sph_chain[slot];
}

template <bool PACKED>
__device__ __forceinline__ SmemScene stage_scene(const DevScene &sc, unsigned char *smem) {
    F4 *nodes = reinterpret_cast<F4 *>(smem);
    F4 *sph = nodes + (size_t)sc.n_nodes * RT_SMEM_NODE_STRIDE / 16; // n_nodes is even: a whole number of F4
    F4 *mats = sph + sc.n_slots;
    F4 *quads = mats + 2 * (size_t)sc.n_mats;
    I2 *meta = reinterpret_cast<I2 *>(quads + (size_t)RT_QUAD_F4 * sc.n_quad_slots);
    uint32_t *chains = reinterpret_cast<uint32_t *>(meta + sc.n_slots + (sc.n_slots & 1));
    uint32_t *sph_chain = chains + (sc.chains ? sc.n_chain_words : 0);
    uint32_t *quad_chain = sph_chain + (sc.chains ? sc.n_slots : 0);
    uint32_t *stack = quad_chain + (sc.chains ? sc.n_quad_slots : 0);
    if (sc.chains) {
        for (uint32_t i = threadIdx.x; i < sc.n_chain_words; i += blockDim.x) chains[i] = __ldg(sc.chains + i);
        for (uint32_t i = threadIdx.x; i < sc.n_slots; i += blockDim.x) sph_chain[i] = __ldg(sc.sph_chain + i);
        for (uint32_t i = threadIdx.x; i < sc.n_quad_slots; i += blockDim.x) quad_chain[i] = __ldg(sc.quad_chain + i);
    }
    const uint4 *src;
    uint4 *dst;
    src = reinterpret_cast<const uint4 *>(sc.nodes), dst = reinterpret_cast<uint4 *>(nodes);
    if constexpr (PACKED) {
    // one thread per pair of sibling nodes: (cL, refL)(hL)(cR, refR)(hR) -> the transposed form box_test_pair reads
    for (uint32_t i = threadIdx.x; i < sc.n_nodes / 2; i += blockDim.x) {
        const uint4 cl = __ldg(src + 4 * i), hl = __ldg(src + 4 * i + 1), cr = __ldg(src + 4 * i + 2), hr = __ldg(src + 4 * i + 3);
        dst[5 * i + 0] = make_uint4(cl.x, cl.y, cr.x, cr.y);
        dst[5 * i + 1] = make_uint4(cl.z, cr.z, hl.z, hr.z);
        dst[5 * i + 2] = make_uint4(hl.x, hl.y, hr.x, hr.y);
        dst[5 * i + 3] = make_uint4(cl.w, cr.w, 0u, 0u);
    }
    } else {
    for (uint32_t i = threadIdx.x; i < 2 * sc.n_nodes; i += blockDim.x) dst[i + (i >> 2)] = __ldg(src + i); // 4 F4 per pair -> 5
    }
    src = reinterpret_cast<const uint4 *>(sc.sph), dst = reinterpret_cast<uint4 *>(sph);
    for (uint32_t i = threadIdx.x; i < sc.n_slots; i += blockDim.x) dst[i] = __ldg(src + i);
    src = reinterpret_cast<const uint4 *>(sc.mats), dst = reinterpret_cast<uint4 *>(mats);
    for (uint32_t i = threadIdx.x; i < 2 * sc.n_mats; i += blockDim.x) dst[i] = __ldg(src + i);
    src = reinterpret_cast<const uint4 *>(sc.quads), dst = reinterpret_cast<uint4 *>(quads);
    for (uint32_t i = threadIdx.x; i < RT_QUAD_F4 * sc.n_quad_slots; i += blockDim.x) dst[i] = __ldg(src + i);
    const uint2 *s2 = reinterpret_cast<const uint2 *>(sc.meta);
    uint2 *d2 = reinterpret_cast<uint2 *>(meta);
    for (uint32_t i = threadIdx.x; i < sc.n_slots; i += blockDim.x) d2[i] = __ldg(s2 + i);
    __syncthreads();
    SmemScene r;
    r.nodes = nodes, r.sph = sph, r.mats = mats, r.quads = quads, r.meta = meta, r.stack = stack;
    r.chains = sc.chains ? chains : nullptr, r.sph_chain = sc.chains ? sph_chain : nullptr, r.quad_chain = quad_chain;
    return r;
}

static size_t smem_total_bytes(const DevScene &s, int block) {
    size_t b = scene_smem_bytes(s) + ((s.n_slots & 1) ? 8 : 0);
    return b + (size_t)s.stack_depth * block * 4;
}

// ---------------------------------------------------------------------------------------------
// render megakernel
// ---------------------------------------------------------------------------------------------
struct RenderParams {
    DevScene sc;
    DevCamera cam;
    uint64_t seed;
    uint32_t pixel_begin;  // first pixel of this pass (row-major index within the call's row set)
    uint32_t row_begin;    // row set of the call (rt_render_opts): image row = row_begin + local row * row_step
    uint32_t row_step;
    uint32_t sample_begin; // global index of the first sample of this pass
    uint32_t spp_pass;     // samples per pixel in this pass
    uint32_t total_paths;  // n_pixels_pass * spp_pass
    float4 *samples;       // [total_paths] radiance of path (pixel - pixel_begin) * spp_pass + k
    unsigned int *counter; // next unclaimed path index
    unsigned long long *stats; // rays, hits, box tests, sphere tests, survivors (RT_N_STATS)
    uint64_t div_spp, div_width; // fast_div multipliers of spp_pass and cam.width (host: fast_div_magic)
    uint32_t regen_min;    // regenerate only when at least this many lanes of the warp are idle
    uint32_t chunk;        // work items a warp claims per atomic (RT_CHUNK by default)
    // two-stage mode (primary_stage_kernel + render_kernel<SPLIT>): paths that survive their first
    // segment, as three float4 arrays of capacity total_paths
    float4 *queue_o;           // (o.xyz, bits path index)
    float4 *queue_d;           // (d.xyz, bits next Philox block | bit 31: radiance parked in the sample slot)
    float4 *queue_t;           // (throughput.xyz, bits hit slot of the first segment: where the next ray starts)
    unsigned int *queue_count; // entries appended by the stage that fills queue_o/d/t
    // stage k > 0 of the staged mode reads the survivors of stage k-1 here (same layout)
    const float4 *in_o, *in_d, *in_t;
    const unsigned int *in_count;
    // per-pixel candidate lists of the call's pixel tile (pixel_candidates_kernel), nullptr = none: RT_LIST_WORDS words per
    // pixel, [count | slots...]; count RT_LIST_OVERFLOW = too many candidates, traverse the tree for that pixel
    const uint32_t *lists;
    int stage_depth; // segments already traced for the paths a stage kernel processes
    size_t queue_stride; // elements per queue array (host-side bookkeeping)
};

// n / d for 32-bit n by one 64x32-bit multiply-high (Lemire & Kaser 2019: M = floor((2^64-1)/d) + 1,
// exact for every 32-bit n and d >= 2; d == 1 is encoded as M == 0).  The path index -> (pixel,
// sample) and pixel -> (row, column) divisions run once per path; a hardware-less 32-bit divide is
// ~16 instructions, this is ~4.
static inline uint64_t fast_div_magic(uint32_t d) { return d <= 1 ? 0ull : 0xFFFFFFFFFFFFFFFFull / d + 1ull; }
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint64_t magic) {
    return magic == 0 ? n : (uint32_t)__umul64hi(magic, (unsigned long long)n);
}

// Pixel `lp` of the call's row set (compact, row-major) -> its index in the full image: the Philox
// counter and the camera ray use the full-image position, so a row set reproduces those rows of the
// full render exactly.  Contiguous rows (row_step 1, every single-GPU render) need no division.
__device__ __forceinline__ uint32_t image_pixel(const RenderParams &p, uint32_t lp) {
    const uint32_t w = (uint32_t)p.cam.width;
    if (p.row_step == 1) return p.row_begin * w + lp;
    // (multiply-high, not `lp / w`: ptxas predicates this path instead of branching around it, and the ~30 instructions
    // of a 32-bit divide were issued — predicated off — for every regenerated path: 1.4 % of the megakernel's instructions)
    const uint32_t jl = fast_div(lp, p.div_width);
    return (p.row_begin + jl * p.row_step) * w + (lp - jl * w);
}

#ifndef RT_GLOBAL_CTA_APPEND
#define RT_GLOBAL_CTA_APPEND 1 /* global-memory scenes: survivors appended once per CTA and round (0: per warp, for A/B:
                                   C4 1368 -> 1265 Msamples/s with the candidate lists, profiles/r02zo — the order of the
                                   queue is what keeps the megakernel's warps on neighbouring parts of the tree) */
#endif
#define RT_CHUNK 256u /* path indices a warp claims per atomic */
#define RT_LIST_WORDS 16u /* one 64-byte line per pixel: count + up to 15 candidate slots */
#define RT_LIST_OVERFLOW 0xFFFFFFFFu
#define RT_MAX_STAGES 8 /* coherent stage kernels before the megakernel (staged mode) */

// The pointers a kernel reads the scene through: global memory, or the CTA's shared-memory copy (stage_scene).
struct SceneView {
    const F4 *nodes, *sph, *mats, *quads;
    const I2 *meta;
    const uint32_t *chains, *sph_chain, *quad_chain;
};
template <bool SMEM, bool QUADS, class Stack, int BLOCK>
__device__ __forceinline__ SceneView open_scene(const DevScene &sc, unsigned char *smem_raw, Stack &stack) {
    SceneView v;
    v.nodes = sc.nodes, v.sph = sc.sph, v.mats = sc.mats, v.quads = sc.quads, v.meta = sc.meta;
    v.chains = sc.chains, v.sph_chain = sc.chains ? sc.sph_chain : nullptr, v.quad_chain = sc.quad_chain;
    if constexpr (SMEM) {
        SmemScene s = stage_scene<RT_PACKED(SMEM, QUADS)>(sc, smem_raw);
        v.nodes = s.nodes, v.sph = s.sph, v.mats = s.mats, v.quads = s.quads, v.meta = s.meta;
        v.chains = s.chains, v.sph_chain = s.sph_chain, v.quad_chain = s.quad_chain;
        stack.base = s.stack + threadIdx.x;
        stack.stride = BLOCK;
    }
    return v;
}

// work counters: warp shuffle reduction, one atomic per warp
template <bool COUNT>
__device__ __forceinline__ void flush_counters(unsigned long long *stats, unsigned long long n_rays, unsigned long long n_hits,
                                               const WorkCounters &wc) {
    unsigned long long v[4] = {n_rays, n_hits, wc.box_tests, wc.sphere_tests};
#pragma unroll
    for (int q = 0; q < (COUNT ? 4 : 2); q++) {
        unsigned long long x = v[q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31u) == 0 && x) atomicAdd(stats + q, x);
    }
}

// The megakernel's loop: one lane = one (pixel, sample) path at a time, dead lanes regenerate from the warp's chunk of
// the pass's work items (all paths, or — SPLIT — the survivors queued by the primary stage).
template <class Stack, bool SMEM, bool COUNT, bool QUADS, bool SPLIT>
__device__ __forceinline__ void path_loop(const RenderParams &p, const SceneView &sv, Stack &stack, unsigned long long &n_rays,
                                          unsigned long long &n_hits, WorkCounters &wc) {
    const F4 *nodes = sv.nodes, *sph = sv.sph, *mats = sv.mats, *quads = sv.quads;
    const I2 *meta = sv.meta;
    const uint32_t *chains = sv.chains, *sph_chain = sv.sph_chain, *quad_chain = sv.quad_chain;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t warp_next = 0, warp_end = 0;
    bool exhausted = false;
    // work items: all paths of the pass, or (two-stage mode) the survivors queued by the primary stage
    const uint32_t total_items = SPLIT ? *p.queue_count : p.total_paths;
    if (SPLIT && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.stats + 4, (unsigned long long)total_items);

    bool alive = false;
    uint32_t idx = 0;
    uint32_t start = RT_REF_NONE; // leaf-start chain of the primitive the current ray leaves (RT_REF_NONE: a camera ray)
    int depth = 0;
    V3 o = v3(0, 0, 0), d = v3(0, 0, 0), thr = v3(1, 1, 1), rad = v3(0, 0, 0);
    PathRng rng;
    rng.init(0, 0, 0);

    for (;;) {
        // ---- regeneration: dead lanes take the next path indices of the warp's chunk ----
        const unsigned dead = __ballot_sync(0xffffffffu, !alive);
        if (dead == 0xffffffffu || (uint32_t)__popc(dead) >= p.regen_min) {
            if (warp_next >= warp_end && !exhausted) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(p.counter, p.chunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= total_items) {
                    exhausted = true;
                } else {
                    warp_next = base;
                    warp_end = min(base + p.chunk, total_items);
                }
            }
            const uint32_t avail = warp_end - warp_next;
            const uint32_t rank = __popc(dead & lt_mask);
            if (!alive && rank < avail) {
                if (SPLIT) { // resume a path after its first segment
                    const uint32_t e = warp_next + rank;
                    RT_DBG(e < RT_DBG_B(queue_cap), RT_DBG_QUEUE);
                    const float4 qo = p.queue_o[e], qd = p.queue_d[e], qt = p.queue_t[e];
                    idx = __float_as_uint(qo.w);
                    RT_DBG(idx < p.total_paths, RT_DBG_SAMPLE);
                    const uint32_t pp = fast_div(idx, p.div_spp), k = idx - pp * p.spp_pass;
                    rng.init(p.seed, image_pixel(p, p.pixel_begin + pp), p.sample_begin + k);
                    rng.block = __float_as_uint(qd.w) & 0x7fffffffu;
                    o = v3(qo.x, qo.y, qo.z), d = v3(qd.x, qd.y, qd.z);
                    thr = v3(qt.x, qt.y, qt.z), rad = v3(0, 0, 0), depth = p.stage_depth;
                    start = chain_of_slot(sph_chain, quad_chain, __float_as_uint(qt.w));
                    if (__float_as_uint(qd.w) & 0x80000000u) { // radiance emitted on the first segment (no reference material does this)
                        const float4 r0 = p.samples[idx];
                        rad = v3(r0.x, r0.y, r0.z);
                    }
                } else {
                    idx = warp_next + rank;
                    const uint32_t pp = fast_div(idx, p.div_spp), k = idx - pp * p.spp_pass;
                    const uint32_t pixel = image_pixel(p, p.pixel_begin + pp);
                    const int j = (int)fast_div(pixel, p.div_width), i = (int)(pixel - (uint32_t)j * p.cam.width);
                    rng.init(p.seed, pixel, p.sample_begin + k);
                    generate_ray(p.cam, rng, i, j, o, d);
                    thr = v3(1, 1, 1), rad = v3(0, 0, 0), depth = 0;
                    start = RT_REF_NONE;
                }
                alive = true;
            }
            warp_next += min(avail, (uint32_t)__popc(dead));
            if (exhausted && __ballot_sync(0xffffffffu, alive) == 0) break;
        }
        if (!alive) continue;

        // ---- one path segment: ray.go:32-54 unrolled front to back ----
        HitRec h;
        trace_closest<Stack, COUNT, QUADS, SMEM, SMEM ? RT_SMEM_NODE_STRIDE : 32, RT_PACKED(SMEM, QUADS)>(nodes, sph, meta, p.sc.root_ref, o, d, 0.001f, INFINITY, stack, h, &wc, quads,
                                                 chains, start);
        n_rays++;
        bool done;
        if (h.slot == RT_REF_NONE) {
            rad = rad + thr * p.cam.background; // ray.go:53
            done = true;
        } else {
            n_hits++;
            V3 atten, emitted;
            bool scattered;
            if (QUADS && (h.slot & RT_HIT_QUAD)) {
                RT_DBG((h.slot & ~RT_HIT_QUAD) < RT_DBG_B(n_quad_slots), RT_DBG_HIT_SLOT);
                const F4 *q = quads + (size_t)RT_QUAD_F4 * (h.slot & ~RT_HIT_QUAD);
                const uint32_t mi = __float_as_uint(q[1].w);
                RT_DBG(mi < RT_DBG_B(n_mats), RT_DBG_MATERIAL);
                const F4 m0 = mats[2 * mi], m1 = mats[2 * mi + 1];
                scattered = shade_hit_quad(m0, m1, p.sc.tex, q, h.t, rng, o, d, atten, emitted);
            } else {
                RT_DBG(h.slot < RT_DBG_B(n_slots), RT_DBG_HIT_SLOT);
                const F4 s = sph[h.slot];
                const int mi = meta[h.slot].y;
                RT_DBG((uint32_t)mi < RT_DBG_B(n_mats), RT_DBG_MATERIAL);
                const F4 m0 = mats[2 * mi], m1 = mats[2 * mi + 1];
                scattered = shade_hit(m0, m1, p.sc.tex, s, h.t, rng, o, d, atten, emitted);
            }
            rad = rad + thr * emitted; // ray.go:41,50
            start = chain_of_slot(sph_chain, quad_chain, h.slot); // the scattered ray leaves this primitive
            if (!scattered) {
                done = true; // ray.go:44-46
            } else {
                thr = thr * atten; // ray.go:48
                depth++;
                done = depth >= p.cam.max_depth; // ray.go:33-35
            }
        }
        if (done) {
            RT_DBG(idx < p.total_paths && idx < RT_DBG_B(samples_cap), RT_DBG_SAMPLE);
            p.samples[idx] = make_float4(rad.x, rad.y, rad.z, 0.0f);
            alive = false;
        }
    }

}

// BLOCK x MINB resident threads per SM bound the register budget (65536 / (BLOCK * MINB)).
template <int BLOCK, int MINB, bool SMEM, bool COUNT, bool QUADS, bool SPLIT = false>
__global__ void __launch_bounds__(BLOCK, MINB) render_kernel(const __grid_constant__ RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename std::conditional<SMEM, StridedStack, LocalStack<RT_LOCAL_STACK>>::type Stack;
    Stack stack;
    const SceneView sv = open_scene<SMEM, QUADS, Stack, BLOCK>(p.sc, smem_raw, stack);
    unsigned long long n_rays = 0, n_hits = 0;
    WorkCounters wc;
    wc.box_tests = wc.sphere_tests = 0;
    path_loop<Stack, SMEM, COUNT, QUADS, SPLIT>(p, sv, stack, n_rays, n_hits, wc);
    flush_counters<COUNT>(p.stats, n_rays, n_hits, wc);
}

// ---------------------------------------------------------------------------------------------
// two-stage mode, stage 1: the first segment of every path, traced as coherent warps
// ---------------------------------------------------------------------------------------------
// Consecutive lanes hold consecutive samples of one pixel, so the 32 primary rays of a warp walk the
// BVH almost identically (measured: ~0.9 lane efficiency against ~0.4 for mixed-depth warps).  Each
// thread generates its camera ray (camera.go:265-299), traces it and applies the first Emit + Scatter
// (ray.go:37-50).  A path that ends here (miss, absorbed, emitter, depth limit) writes its radiance;
// a survivor is appended to the queue (warp-aggregated atomic) and finished by render_kernel<SPLIT>.
// Same functions, same order per path as the one-stage kernel: the image is bit-identical.
// FIRST = false: the same for a later segment — the rays come from the previous stage's queue (in
// pixel-neighbour order, all lanes alive) instead of the camera.
// One round of the primary stage for the calling thread: work item `item` of `n_items` (lanes past the end idle, but
// take part in the warp- / CTA-wide queue append).  cta_count / cta_base: BLOCK / 32 words of shared memory (global-memory
// scenes append once per CTA and round, behind two barriers — every thread of the CTA must then call this together).
template <class Stack, int BLOCK, bool SMEM, bool COUNT, bool QUADS, bool FIRST>
__device__ __forceinline__ void primary_round(const RenderParams &p, const SceneView &sv, Stack &stack, uint32_t item, uint32_t n_items,
                                              unsigned long long &n_rays, unsigned long long &n_hits, WorkCounters &wc,
                                              uint32_t *cta_count, uint32_t *cta_base) {
    const F4 *nodes = sv.nodes, *sph = sv.sph, *mats = sv.mats, *quads = sv.quads;
    const I2 *meta = sv.meta;
    const uint32_t *chains = sv.chains, *sph_chain = sv.sph_chain, *quad_chain = sv.quad_chain;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t idx = item;
    bool survive = false, carries = false;
    V3 o = v3(0, 0, 0), d = v3(0, 0, 0), thr = v3(1, 1, 1);
    uint32_t block = 0, hit_slot = RT_REF_NONE;
    if (item < n_items) {
        PathRng rng;
        V3 rad = v3(0, 0, 0);
        uint32_t start = RT_REF_NONE;
        if (FIRST) {
            const uint32_t pp = fast_div(idx, p.div_spp), k = idx - pp * p.spp_pass;
            const uint32_t pixel = image_pixel(p, p.pixel_begin + pp);
            const int j = (int)fast_div(pixel, p.div_width), i = (int)(pixel - (uint32_t)j * p.cam.width);
            rng.init(p.seed, pixel, p.sample_begin + k);
            generate_ray(p.cam, rng, i, j, o, d);
        } else {
            RT_DBG(item < RT_DBG_B(queue_cap), RT_DBG_QUEUE);
            const float4 qo = p.in_o[item], qd = p.in_d[item], qt = p.in_t[item];
            idx = __float_as_uint(qo.w);
            const uint32_t pp = fast_div(idx, p.div_spp), k = idx - pp * p.spp_pass;
            rng.init(p.seed, image_pixel(p, p.pixel_begin + pp), p.sample_begin + k);
            rng.block = __float_as_uint(qd.w) & 0x7fffffffu;
            o = v3(qo.x, qo.y, qo.z), d = v3(qd.x, qd.y, qd.z), thr = v3(qt.x, qt.y, qt.z);
            start = chain_of_slot(sph_chain, quad_chain, __float_as_uint(qt.w));
            if (__float_as_uint(qd.w) & 0x80000000u) { // radiance parked by an earlier stage
                const float4 r0 = p.samples[idx];
                rad = v3(r0.x, r0.y, r0.z);
            }
        }
        HitRec h;
        const uint32_t *list = nullptr;
        uint32_t n_list = RT_LIST_OVERFLOW;
        if (FIRST && p.lists != nullptr) { // the candidates of this path's pixel (camera rays only)
            list = p.lists + (size_t)fast_div(idx, p.div_spp) * RT_LIST_WORDS;
            n_list = list[0];
            RT_DBG(n_list == RT_LIST_OVERFLOW || n_list < RT_LIST_WORDS, RT_DBG_LEAF);
        }
        if (FIRST && n_list != RT_LIST_OVERFLOW)
            trace_candidates<COUNT, QUADS>(list + 1, n_list, sph, meta, quads, o, d, 0.001f, INFINITY, h, &wc);
        else
            trace_closest<Stack, COUNT, QUADS, false, SMEM ? RT_SMEM_NODE_STRIDE : 32, RT_PACKED(SMEM, QUADS)>(nodes, sph, meta, p.sc.root_ref, o, d, 0.001f, INFINITY, stack, h, &wc, quads,
                                           FIRST ? nullptr : chains, start);
        n_rays++;
        hit_slot = h.slot;
        if (h.slot == RT_REF_NONE) {
            rad = rad + thr * p.cam.background; // ray.go:53
        } else {
            n_hits++;
            V3 atten, emitted;
            bool scattered;
            if (QUADS && (h.slot & RT_HIT_QUAD)) {
                RT_DBG((h.slot & ~RT_HIT_QUAD) < RT_DBG_B(n_quad_slots), RT_DBG_HIT_SLOT);
                const F4 *q = quads + (size_t)RT_QUAD_F4 * (h.slot & ~RT_HIT_QUAD);
                const uint32_t mi = __float_as_uint(q[1].w);
                RT_DBG(mi < RT_DBG_B(n_mats), RT_DBG_MATERIAL);
                scattered = shade_hit_quad(mats[2 * mi], mats[2 * mi + 1], p.sc.tex, q, h.t, rng, o, d, atten, emitted);
            } else {
                RT_DBG(h.slot < RT_DBG_B(n_slots), RT_DBG_HIT_SLOT);
                const F4 s = sph[h.slot];
                const int mi = meta[h.slot].y;
                RT_DBG((uint32_t)mi < RT_DBG_B(n_mats), RT_DBG_MATERIAL);
                scattered = shade_hit(mats[2 * mi], mats[2 * mi + 1], p.sc.tex, s, h.t, rng, o, d, atten, emitted);
            }
            rad = rad + thr * emitted; // ray.go:41,50
            if (scattered) {
                thr = thr * atten; // ray.go:48
                survive = p.stage_depth + 1 < p.cam.max_depth; // ray.go:33-35
            }
        }
        block = rng.block;
        // no material of the reference both emits and scatters, so a survivor normally carries no
        // radiance; if one ever does, it is parked in the sample slot and flagged in the queue entry
        carries = survive && (rad.x != 0.0f || rad.y != 0.0f || rad.z != 0.0f);
        RT_DBG(idx < p.total_paths && idx < RT_DBG_B(samples_cap), RT_DBG_SAMPLE);
        if (!survive || carries) p.samples[idx] = make_float4(rad.x, rad.y, rad.z, 0.0f);
    }
    // append survivors, consecutive entries for consecutive lanes
    const unsigned m = __ballot_sync(0xffffffffu, survive);
    uint32_t first_entry = 0;
    if constexpr (SMEM || !RT_GLOBAL_CTA_APPEND) { // one atomic per warp
        if (m) {
            if (lane == (unsigned)(__ffs(m) - 1)) first_entry = atomicAdd(p.queue_count, (unsigned)__popc(m));
            first_entry = __shfl_sync(0xffffffffu, first_entry, __ffs(m) - 1);
        }
    } else {
        // Scene in global memory (1 M spheres): one atomic per CTA and round, behind two barriers.  The
        // barriers keep the CTA's warps on neighbouring pixels, which the L1-resident top of the tree likes
        // (C4 +5.5 %); with the scene in shared memory they only add waiting (C2 -1 %), hence the split.
        // The round count is the same for every thread of the grid, so the barriers are reached by all.
        const unsigned warp = threadIdx.x >> 5;
        if (lane == 0) cta_count[warp] = (uint32_t)__popc(m);
        __syncthreads();
        if (warp == 0) { // exclusive prefix over the warps' counts, then the CTA's reservation
            const uint32_t c = lane < BLOCK / 32 ? cta_count[lane] : 0u;
            uint32_t incl = c;
#pragma unroll
            for (int off = 1; off < BLOCK / 32; off <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)lane >= off) incl += up;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, BLOCK / 32 - 1);
            uint32_t base = 0;
            if (lane == 0 && total) base = atomicAdd(p.queue_count, total);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lane < BLOCK / 32) cta_base[lane] = base + incl - c;
        }
        __syncthreads();
        first_entry = cta_base[warp];
    }
    if (survive) {
        const uint32_t e = first_entry + __popc(m & ((1u << lane) - 1u));
        RT_DBG(e < p.total_paths && e < RT_DBG_B(queue_cap), RT_DBG_QUEUE);
        p.queue_o[e] = make_float4(o.x, o.y, o.z, __uint_as_float(idx));
        p.queue_d[e] = make_float4(d.x, d.y, d.z, __uint_as_float(block | (carries ? 0x80000000u : 0u)));
        p.queue_t[e] = make_float4(thr.x, thr.y, thr.z, __uint_as_float(hit_slot)); // the primitive the survivor leaves
    }
}

template <int BLOCK, bool SMEM, bool COUNT, bool QUADS, bool FIRST = true>
__global__ void __launch_bounds__(BLOCK) primary_stage_kernel(const __grid_constant__ RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename std::conditional<SMEM, StridedStack, LocalStack<RT_LOCAL_STACK>>::type Stack;
    Stack stack;
    const SceneView sv = open_scene<SMEM, QUADS, Stack, BLOCK>(p.sc, smem_raw, stack);
    __shared__ uint32_t cta_count[BLOCK / 32], cta_base[BLOCK / 32];
    unsigned long long n_rays = 0, n_hits = 0;
    WorkCounters wc;
    wc.box_tests = wc.sphere_tests = 0;
    // whole warps iterate together (the trip count is warp-uniform), lanes past the end idle
    const uint32_t n_items = FIRST ? p.total_paths : *p.in_count;
    const uint32_t n_rounds = (n_items + BLOCK * gridDim.x - 1) / (BLOCK * gridDim.x);
    for (uint32_t r = 0; r < n_rounds; r++)
        primary_round<Stack, BLOCK, SMEM, COUNT, QUADS, FIRST>(p, sv, stack, (r * gridDim.x + blockIdx.x) * BLOCK + threadIdx.x, n_items,
                                                               n_rays, n_hits, wc, cta_count, cta_base);
    flush_counters<COUNT>(p.stats, n_rays, n_hits, wc);
}

// ---------------------------------------------------------------------------------------------
// per-pixel candidate lists (primary stage)
// ---------------------------------------------------------------------------------------------
// Every camera ray of a pixel — all samples of all passes — starts on the defocus disk and passes through the pixel's
// footprint, so the primitives ANY of them can hit are few (C2: 4.3 on average, profiles/experiments/r01_pixel_beam_*).
// One thread per pixel walks the BVH once with the pixel's beam (rt_trace.h: beam_box_test, conservative interval
// arithmetic) and writes the slots of the leaves it reaches; the primary stage then tests a path's camera ray against
// that list (trace_candidates) instead of traversing the tree 500 times per pixel.  The closest hit is an argmin over
// all primitives, hence also over any superset of the ones a ray can hit: results are unchanged, bit for bit.
// A pixel with more than RT_LIST_WORDS-1 candidates is marked RT_LIST_OVERFLOW and keeps the traversal.
template <bool QUADS>
__global__ void __launch_bounds__(128) pixel_candidates_kernel(const __grid_constant__ RenderParams p, uint32_t n_pixels,
                                                               uint32_t *__restrict__ lists) {
    const uint32_t pp = blockIdx.x * blockDim.x + threadIdx.x;
    if (pp >= n_pixels) return;
    const uint32_t pixel = image_pixel(p, p.pixel_begin + pp);
    const int j = (int)(pixel / (uint32_t)p.cam.width), i = (int)(pixel - (uint32_t)j * p.cam.width);
    const Beam b = pixel_beam(p.cam, i, j);
    uint32_t *out = lists + (size_t)pp * RT_LIST_WORDS;
    const uint32_t n = beam_candidates<QUADS>(p.sc.nodes, p.sc.root_ref, b, out + 1, RT_LIST_WORDS - 1);
    out[0] = n > RT_LIST_WORDS - 1 ? RT_LIST_OVERFLOW : n;
}

// accum[pixel] (+)= sum_k samples[(pixel - pixel_begin) * spp_pass + k], k ascending: the FP32
// summation order of camera.go:255-260.  first_pass: start from zero instead of accum.
__global__ void reduce_kernel(const float4 *__restrict__ samples, float *__restrict__ accum, uint32_t pixel_begin,
                              uint32_t n_pixels, uint32_t spp_pass, int first_pass) {
    const uint32_t pp = blockIdx.x * blockDim.x + threadIdx.x;
    if (pp >= n_pixels) return;
    RT_DBG((unsigned long long)pixel_begin + pp < RT_DBG_B(n_pixels), RT_DBG_PIXEL);
    RT_DBG(((size_t)pp + 1) * spp_pass <= RT_DBG_B(samples_cap), RT_DBG_SAMPLE);
    float *a = accum + 3 * (size_t)(pixel_begin + pp);
    V3 sum = first_pass ? v3(0, 0, 0) : v3(a[0], a[1], a[2]);
    const float4 *s = samples + (size_t)pp * spp_pass;
    for (uint32_t k = 0; k < spp_pass; k++) {
        const float4 r = s[k];
        sum = sum + v3(r.x, r.y, r.z);
    }
    a[0] = sum.x, a[1] = sum.y, a[2] = sum.z;
}

__global__ void resolve_kernel(const float *__restrict__ accum, uint8_t *__restrict__ rgb, uint32_t n_pixels,
                               float inv_spp) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= n_pixels) return;
    const float *a = accum + 3 * (size_t)pix;
    uint8_t out[3];
    resolve_pixel(v3(a[0], a[1], a[2]), inv_spp, out);
    rgb[3 * (size_t)pix + 0] = out[0], rgb[3 * (size_t)pix + 1] = out[1], rgb[3 * (size_t)pix + 2] = out[2];
}

// ---------------------------------------------------------------------------------------------
// parity hooks
// ---------------------------------------------------------------------------------------------
template <int BLOCK, bool SMEM, bool QUADS>
__global__ void __launch_bounds__(BLOCK) trace_kernel(const __grid_constant__ DevScene sc, const float *__restrict__ origins,
                                                      const float *__restrict__ dirs, long long n, float tmin, float tmax,
                                                      int32_t *__restrict__ id_out, float *__restrict__ t_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const F4 *nodes = sc.nodes, *sph = sc.sph, *quads = sc.quads;
    const I2 *meta = sc.meta;
    typedef typename std::conditional<SMEM, StridedStack, LocalStack<RT_LOCAL_STACK>>::type Stack;
    Stack stack;
    if constexpr (SMEM) {
        SmemScene s = stage_scene<RT_PACKED(SMEM, QUADS)>(sc, smem_raw);
        nodes = s.nodes, sph = s.sph, quads = s.quads, meta = s.meta;
        stack.base = s.stack + threadIdx.x;
        stack.stride = BLOCK;
    }
    for (long long i = (long long)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * BLOCK) {
        const V3 o = v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        const V3 d = v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        HitRec h;
        trace_closest<Stack, false, QUADS, false, SMEM ? RT_SMEM_NODE_STRIDE : 32, RT_PACKED(SMEM, QUADS)>(nodes, sph, meta, sc.root_ref, o, d, tmin, tmax, stack, h, nullptr, quads);
        if (h.slot == RT_REF_NONE) {
            id_out[i] = -1, t_out[i] = 0.0f;
        } else {
            id_out[i] = slot_object_id(h.slot, meta, quads), t_out[i] = h.t;
        }
    }
}

__global__ void primary_kernel(DevCamera cam, uint64_t seed, uint32_t pixel_begin, uint32_t sample_begin,
                               uint32_t spp, long long n, float *__restrict__ origins, float *__restrict__ dirs) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const uint32_t pp = (uint32_t)(idx / spp), k = (uint32_t)(idx - (long long)pp * spp);
    const uint32_t pixel = pixel_begin + pp;
    const int j = (int)(pixel / (uint32_t)cam.width), i = (int)(pixel - (uint32_t)j * cam.width);
    PathRng rng;
    rng.init(seed, pixel, sample_begin + k);
    V3 o, d;
    generate_ray(cam, rng, i, j, o, d);
    origins[3 * idx] = o.x, origins[3 * idx + 1] = o.y, origins[3 * idx + 2] = o.z;
    dirs[3 * idx] = d.x, dirs[3 * idx + 1] = d.y, dirs[3 * idx + 2] = d.z;
}


// dst += src (multi-GPU gather of FP32 accumulators, in a fixed device order)
__global__ void add_kernel(float *__restrict__ dst, const float *__restrict__ src, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = dst[i] + src[i];
}

#endif // RT_KERNELS_CUH
