// bvh_device.cuh — the BVH of a large sphere scene built ON the GPU (replaces NewBVH, bvh.go:142-185, for scenes
// whose host build would cost more than rendering them: config C4's 1 M spheres took ~120 ms of 16 host cores per
// rt_scene_create against ~95 ms of rendering).
//
// The reference's tree is a random-axis median split rebuilt on every run; its topology is not part of any result
// (the closest hit is an argmin over all primitives, any tree + conservative culling returns it),
// so the device build is free to be a linear BVH:
//   1. bd_prim_boxes   padded box of every sphere — the SAME padding rule as the host builder (bvh_build.cpp), in
//                      float64 — and the bounds of the centres;
//   2. bd_keys         63-bit Morton code of the centre; spheres whose diameter exceeds a quarter of the scene
//                      (the r = 1000 "ground") get the largest key: they sort to the end and are attached above
//                      the tree instead of inflating every box on their Morton path;
//   3. cub::DeviceRadixSort (library primitive; the only one — everything else is kernels of this file);
//   4. bd_karras       the radix tree over the sorted codes, one thread per inner node (Karras 2012), ties between
//                      equal codes broken by position;
//   5. bd_fit          boxes bottom-up, second arrival at a node continues (one atomic counter per node);
//   6. bd_live + scan  subtrees of at most `max_leaf` spheres collapse into one leaf (their slots are contiguous in
//                      Morton order); the surviving inner nodes are numbered by a prefix sum;
//   7. bd_emit, bd_top the device layout the traversal kernels read (rt_trace.h): pairs of sibling nodes as
//                      (centre, ref)(half-extent), widened exactly as make_device_nodes does on the host; a short
//                      chain of pairs at the top holds the huge spheres;
//   8. bd_slots        sphere and (object ID, material) records in slot order;  bd_depth: the depth the traversal
//                      stack must hold.
// Nothing is read back but one small statistics record.  A larger origin radius (rt_render from a far camera) runs
// the same build again with the new padding: same codes, same topology, a few milliseconds.
#ifndef RT_BVH_DEVICE_CUH
#define RT_BVH_DEVICE_CUH

#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "rt_trace.h"

#define BD_MAX_LARGE 64 /* huge spheres attached above the tree; more than that: host build */

struct BdStats {
    int cb_lo[3], cb_hi[3];   // bounds of the sphere centres (float bits, order-preserving encoding)
    unsigned int n_large;     // spheres attached above the tree
    unsigned int n_live;      // inner nodes that survive the leaf collapse
    unsigned int max_depth;   // deepest chain of inner nodes, top chain included
    unsigned int pad_min, pad_max; // float bits (pads are positive: unsigned order = float order)
    unsigned int surface_extent;   // float bits: max over spheres of (|c - m| - |r|), clamped at 0
    unsigned int root_ref;
    unsigned int n_pairs;     // pairs of device nodes written (top chain + live inner nodes)
    unsigned int large[BD_MAX_LARGE]; // their sorted positions are n - n_large ..; (unused, kept for debugging)
};

__device__ __forceinline__ int bd_enc(float f) { // order-preserving float -> int
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float bd_dec(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct BdParams {
    const rt_sphere *raw; // caller order
    const uint32_t *ids;  // object IDs (nullptr: the index)
    uint32_t n;
    double m[3];          // the scene's centre (host: sample median)
    double origin_radius;
    int max_leaf;
    BdStats *st;
    // per sphere, caller order
    float4 *lo, *hi;      // padded box
    // sorted order
    unsigned long long *keys, *keys_sorted;
    uint32_t *idx, *idx_sorted;
    // radix tree over the nn = n - n_large ordinary spheres: inner node i in [0, nn-1)
    int *left, *right;    // child: >= 0 inner node, < 0: leaf ~child (sorted position)
    int *parent;          // [0, nn-1): of inner nodes; [n, n + nn): of leaves (offset n)
    uint32_t *first, *last;
    float4 *nlo, *nhi;    // inner node boxes
    unsigned int *arrive;
    uint32_t *live, *compact;
    // outputs
    F4 *nodes; // pairs
    F4 *sph;
    I2 *meta;
};

__global__ void bd_init(BdStats *st) {
    for (int k = 0; k < 3; k++) st->cb_lo[k] = 0x7fffffff, st->cb_hi[k] = (int)0x80000000;
    st->n_large = st->n_live = st->max_depth = 0;
    st->pad_min = 0x7f800000u, st->pad_max = 0u, st->surface_extent = 0u, st->root_ref = RT_REF_NONE, st->n_pairs = 0;
}

// The padding rule of bvh_build.cpp (padded_box), float64.  Scene statistics: one atomic per warp and quantity.
__global__ void bd_prim_boxes(BdParams p) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = g < p.n;
    int clo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, chi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    unsigned int pmin = 0x7f800000u, pmax = 0u, sext = 0u;
    if (in) {
        const double U = 1.0 / 16777216.0, K_DISC = 24.0;
        const rt_sphere s = p.raw[g];
        const double r = fabs((double)s.r), D = p.origin_radius + r;
        const double cmax = fmax(fabs((double)s.cx), fmax(fabs((double)s.cy), fabs((double)s.cz)));
        const double mmax = fmax(fabs(p.m[0]), fmax(fabs(p.m[1]), fabs(p.m[2])));
        double pad = K_DISC * U * D * D / (2.0 * fmax(r, 1e-30)) + 8.0 * U * (D + mmax + cmax + r);
        pad = fmin(pad, D + r);
        const float cc[3] = {s.cx, s.cy, s.cz};
        float lo[3], hi[3];
        for (int k = 0; k < 3; k++) {
            lo[k] = nextafterf((float)((double)cc[k] - r - pad), -INFINITY);
            hi[k] = nextafterf((float)((double)cc[k] + r + pad), INFINITY);
            clo[k] = chi[k] = bd_enc(cc[k]);
        }
        p.lo[g] = make_float4(lo[0], lo[1], lo[2], 0.0f), p.hi[g] = make_float4(hi[0], hi[1], hi[2], 0.0f);
        pmin = pmax = __float_as_uint((float)pad);
        const double dx = s.cx - p.m[0], dy = s.cy - p.m[1], dz = s.cz - p.m[2];
        sext = __float_as_uint((float)fmax(0.0, sqrt(dx * dx + dy * dy + dz * dz) - r));
    }
    for (int k = 0; k < 3; k++) clo[k] = __reduce_min_sync(0xffffffffu, clo[k]), chi[k] = __reduce_max_sync(0xffffffffu, chi[k]);
    pmin = __reduce_min_sync(0xffffffffu, pmin), pmax = __reduce_max_sync(0xffffffffu, pmax);
    sext = __reduce_max_sync(0xffffffffu, sext);
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; k++) atomicMin(&p.st->cb_lo[k], clo[k]), atomicMax(&p.st->cb_hi[k], chi[k]);
        atomicMin(&p.st->pad_min, pmin), atomicMax(&p.st->pad_max, pmax), atomicMax(&p.st->surface_extent, sext);
    }
}

__device__ __forceinline__ unsigned long long bd_spread21(unsigned long long v) { // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void bd_keys(BdParams p) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n) return;
    const rt_sphere s = p.raw[g];
    float lo[3], ext[3], maxext = 0.0f;
    for (int k = 0; k < 3; k++) {
        lo[k] = bd_dec(p.st->cb_lo[k]);
        ext[k] = bd_dec(p.st->cb_hi[k]) - lo[k];
        maxext = fmaxf(maxext, ext[k]);
    }
    p.idx[g] = g;
    // a sphere as large as a quarter of the whole scene would put its box around every node on its Morton path
    if (p.n > 1 && 2.0f * fabsf(s.r) > 0.25f * maxext && maxext > 0.0f) {
        atomicAdd(&p.st->n_large, 1u);
        p.keys[g] = ~0ull;
        return;
    }
    const float c[3] = {s.cx, s.cy, s.cz};
    unsigned long long q[3];
    for (int k = 0; k < 3; k++) {
        const float t = ext[k] > 0.0f ? (c[k] - lo[k]) / ext[k] : 0.0f;
        q[k] = (unsigned long long)fminf(fmaxf(t * 2097152.0f, 0.0f), 2097151.0f);
    }
    p.keys[g] = bd_spread21(q[0]) << 2 | bd_spread21(q[1]) << 1 | bd_spread21(q[2]); // < 2^63: below the "large" key
}

// number of leading bits positions i and j of the sorted sequence share (Karras 2012), -1 outside [0, nn)
__device__ __forceinline__ int bd_delta(const unsigned long long *keys, int nn, int i, int j) {
    if (j < 0 || j >= nn) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(i ^ j) : __clzll((long long)(a ^ b));
}

__global__ void bd_karras(BdParams p) {
    const int nn = (int)(p.n - p.st->n_large);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn - 1) return;
    const unsigned long long *keys = p.keys_sorted;
    const int d = bd_delta(keys, nn, i, i + 1) - bd_delta(keys, nn, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = bd_delta(keys, nn, i, i - d);
    int lmax = 2;
    while (bd_delta(keys, nn, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (bd_delta(keys, nn, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = bd_delta(keys, nn, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (bd_delta(keys, nn, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = lo == gamma ? ~gamma : gamma, rc = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    p.left[i] = lc, p.right[i] = rc;
    p.first[i] = (uint32_t)lo, p.last[i] = (uint32_t)hi;
    if (lc >= 0) p.parent[lc] = i;
    else p.parent[p.n + (uint32_t)~lc] = i;
    if (rc >= 0) p.parent[rc] = i;
    else p.parent[p.n + (uint32_t)~rc] = i;
    if (i == 0) p.parent[0] = -1;
    p.arrive[i] = 0;
}

// box of child c of an inner node, read past L1 (it was written by another thread, possibly on another SM)
__device__ __forceinline__ void bd_child_box(const BdParams &p, int c, float4 *lo, float4 *hi) {
    if (c < 0) {
        const uint32_t g = p.idx_sorted[(uint32_t)~c];
        *lo = p.lo[g], *hi = p.hi[g];
    } else {
        *lo = __ldcg(&p.nlo[c]), *hi = __ldcg(&p.nhi[c]);
    }
}

// Bottom-up: one thread per leaf climbs; the first thread to reach a node stops, the second — both children are
// complete then — writes the union and goes on (Karras 2012).  No thread ever waits for another.
__global__ void bd_fit(BdParams p) {
    const int nn = (int)(p.n - p.st->n_large);
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nn || nn < 2) return;
    int node = p.parent[p.n + (uint32_t)j];
    while (node >= 0) {
        __threadfence(); // the box this thread wrote one level below is visible before it announces itself
        if (atomicAdd(&p.arrive[node], 1u) == 0) return;
        float4 llo, lhi, rlo, rhi;
        bd_child_box(p, p.left[node], &llo, &lhi);
        bd_child_box(p, p.right[node], &rlo, &rhi);
        p.nlo[node] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
        p.nhi[node] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
        node = p.parent[node];
    }
}

__global__ void bd_live(BdParams p) {
    const int nn = (int)(p.n - p.st->n_large);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int)p.n) return;
    p.live[i] = (i < nn - 1 && p.last[i] - p.first[i] + 1 > (uint32_t)p.max_leaf) ? 1u : 0u;
}

// (min, max) -> (centre, half-extent), widened as make_device_nodes (bvh_build.cpp) widens them on the host
__device__ __forceinline__ void bd_centre_form(float4 lo, float4 hi, uint32_t ref, F4 *out) {
    const double U = 5.960464477539063e-08;
    const float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
    float c[3], e[3];
    for (int k = 0; k < 3; k++) {
        c[k] = (float)(0.5 * ((double)l[k] + (double)h[k]));
        double half = fmax((double)h[k] - (double)c[k], (double)c[k] - (double)l[k]);
        half += 4.0 * U * (fabs((double)c[k]) + half);
        e[k] = nextafterf((float)half, INFINITY);
        if (!((double)c[k] - (double)e[k] <= (double)l[k] && (double)c[k] + (double)e[k] >= (double)h[k])) e[k] = INFINITY;
    }
    out[0].x = c[0], out[0].y = c[1], out[0].z = c[2], out[0].w = __uint_as_float(ref);
    out[1].x = e[0], out[1].y = e[1], out[1].z = e[2], out[1].w = 0.0f;
}

// reference + box of child `c` of a live inner node
__device__ __forceinline__ void bd_child(const BdParams &p, int c, uint32_t n_top, uint32_t *ref, float4 *lo, float4 *hi) {
    if (c < 0) { // one sphere
        const uint32_t j = (uint32_t)~c, g = p.idx_sorted[j];
        *ref = RT_LEAF | j << 3, *lo = p.lo[g], *hi = p.hi[g];
    } else {
        *lo = p.nlo[c], *hi = p.nhi[c];
        if (p.live[c]) *ref = 2u * (n_top + p.compact[c]);
        else *ref = RT_LEAF | p.first[c] << 3 | (p.last[c] - p.first[c]); // collapsed subtree: one leaf over its range
    }
}

__global__ void bd_emit(BdParams p) {
    const int nn = (int)(p.n - p.st->n_large);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn - 1 || !p.live[i]) return;
    const uint32_t n_top = p.st->n_large;
    uint32_t lref, rref;
    float4 llo, lhi, rlo, rhi;
    bd_child(p, p.left[i], n_top, &lref, &llo, &lhi);
    bd_child(p, p.right[i], n_top, &rref, &rlo, &rhi);
    F4 *out = p.nodes + 4 * (size_t)(n_top + p.compact[i]);
    bd_centre_form(llo, lhi, lref, out);
    bd_centre_form(rlo, rhi, rref, out + 2);
}

// The top of the tree, one thread: the huge spheres (sorted positions nn .. n-1) hang off a chain of pairs
//   pair k = (leaf of huge sphere k, the rest: pair k+1 ... and finally the ordinary tree).
__global__ void bd_top(BdParams p) {
    if (blockIdx.x || threadIdx.x) return;
    BdStats *st = p.st;
    const uint32_t L = st->n_large, nn = p.n - L;
    st->n_live = nn >= 2 ? p.compact[nn - 2] + p.live[nn - 2] : 0; // exclusive scan + last flag
    // the ordinary tree's root reference and box
    uint32_t r0 = RT_REF_NONE;
    float4 lo0 = make_float4(0, 0, 0, 0), hi0 = make_float4(-1e30f, -1e30f, -1e30f, 0); // never hit
    if (nn == 1) {
        const uint32_t g = p.idx_sorted[0];
        r0 = RT_LEAF, lo0 = p.lo[g], hi0 = p.hi[g]; // slot 0, one sphere
    } else if (nn >= 2) {
        lo0 = p.nlo[0], hi0 = p.nhi[0];
        r0 = p.live[0] ? 2u * (L + p.compact[0]) : (RT_LEAF | (nn - 1));
    }
    st->n_pairs = L + st->n_live;
    if (L == 0) { // the ordinary tree is the whole tree (its root may be a single leaf: trace_closest starts there)
        st->root_ref = r0;
        return;
    }
    if (L > BD_MAX_LARGE) return; // the host falls back to its own builder
    // boxes of "everything after huge sphere k", from the end
    float4 rest_lo = lo0, rest_hi = hi0;
    uint32_t rest_ref = r0;
    for (int k = (int)L - 1; k >= 0; k--) {
        const uint32_t j = nn + (uint32_t)k, g = p.idx_sorted[j];
        F4 *out = p.nodes + 4 * (size_t)k;
        bd_centre_form(p.lo[g], p.hi[g], RT_LEAF | j << 3, out);
        if (rest_ref == RT_REF_NONE) { // no ordinary spheres at all: a box that is never hit
            out[2] = out[0], out[3] = out[1];
            out[2].x = out[2].y = out[2].z = 0.0f, out[3].x = out[3].y = out[3].z = -1e30f;
        } else {
            bd_centre_form(rest_lo, rest_hi, rest_ref, out + 2);
        }
        const float4 blo = p.lo[g], bhi = p.hi[g];
        if (rest_ref == RT_REF_NONE) rest_lo = blo, rest_hi = bhi;
        else {
            rest_lo = make_float4(fminf(rest_lo.x, blo.x), fminf(rest_lo.y, blo.y), fminf(rest_lo.z, blo.z), 0);
            rest_hi = make_float4(fmaxf(rest_hi.x, bhi.x), fmaxf(rest_hi.y, bhi.y), fmaxf(rest_hi.z, bhi.z), 0);
        }
        rest_ref = 2u * (uint32_t)k;
    }
    st->root_ref = 0;
}

__global__ void bd_slots(BdParams p) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.n) return;
    const uint32_t g = p.idx_sorted[j];
    const rt_sphere s = p.raw[g];
    F4 v;
    v.x = s.cx, v.y = s.cy, v.z = s.cz, v.w = s.r;
    p.sph[j] = v;
    I2 m;
    m.x = (int32_t)(p.ids ? p.ids[g] : g), m.y = (int32_t)s.material;
    p.meta[j] = m;
}

__global__ void bd_depth(BdParams p) {
    const int nn = (int)(p.n - p.st->n_large);
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int depth = 0; // lanes past the end stay in the warp for the reduction below
    if (j < nn) {
        depth = p.st->n_large;
        if (nn >= 2)
            for (int node = p.parent[p.n + (uint32_t)j]; node >= 0; node = p.parent[node]) depth += p.live[node];
    }
    depth = __reduce_max_sync(0xffffffffu, depth); // one atomic per warp
    if ((threadIdx.x & 31) == 0 && depth) atomicMax(&p.st->max_depth, depth);
}

// pack_materials (bvh_build.cpp) on the device: the texture folded into each 32-byte material record (rt_shade.h)
__global__ void bd_pack_materials(const rt_material *__restrict__ mats, const rt_texture *__restrict__ tex, uint32_t n, F4 *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rt_material m = mats[i];
    F4 m0 = {0, 0, 0, 0}, m1 = {0, 0, 0, 0};
    uint32_t code = RT_CODE(m.kind, 0, 0);
    if (m.kind == RT_MAT_METAL) {
        m0.x = m.albedo[0], m0.y = m.albedo[1], m0.z = m.albedo[2], m0.w = m.fuzz;
    } else if (m.kind == RT_MAT_DIELECTRIC) {
        m0.w = m.ior;
        m1.x = __fdiv_rn(1.0f, m.ior); // materials.go:94
    } else {
        const rt_texture t = tex[m.texture];
        code = RT_CODE(m.kind, t.kind, t.kind == RT_TEX_IMAGE ? t.image : 0);
        if (t.kind == RT_TEX_CHECKER) {
            m0.x = t.a[0], m0.y = t.a[1], m0.z = t.a[2];
            m0.w = __fdiv_rn(1.0f, t.scale); // materials.go:128
            m1.x = t.b[0], m1.y = t.b[1], m1.z = t.b[2];
        } else if (t.kind == RT_TEX_NOISE) {
            code = RT_CODE(m.kind, t.kind, t.image);
            m0.w = t.scale;
        } else if (t.kind == RT_TEX_IMAGE) {
            m0.x = t.oob[0], m0.y = t.oob[1], m0.z = t.oob[2];
        } else {
            m0.x = t.a[0], m0.y = t.a[1], m0.z = t.a[2];
        }
    }
    m1.w = __uint_as_float(code);
    out[2 * (size_t)i] = m0, out[2 * (size_t)i + 1] = m1;
}

// Host driver: `raw` / `ids` are device copies of the scene's spheres; nodes (>= 4*n F4), sph, meta (>= n) are device
// buffers of the scene handle.  Returns a cudaError_t; *out receives the statistics (root_ref, n_pairs, max_depth ...).
static cudaError_t device_build_bvh(const rt_sphere *raw, const uint32_t *ids, uint32_t n, const double m[3], double origin_radius,
                                    int max_leaf, F4 *nodes, F4 *sph, I2 *meta, cudaStream_t st, BdStats *out) {
    BdParams p;
    memset(&p, 0, sizeof p);
    p.raw = raw, p.ids = ids, p.n = n, p.origin_radius = origin_radius, p.max_leaf = max_leaf;
    p.m[0] = m[0], p.m[1] = m[1], p.m[2] = m[2];
    p.nodes = nodes, p.sph = sph, p.meta = meta;
    // one temporary arena from the stream's pool
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    const size_t N = n;
    const size_t o_st = take(sizeof(BdStats)), o_lo = take(N * 16), o_hi = take(N * 16), o_k = take(N * 8), o_ks = take(N * 8);
    const size_t o_i = take(N * 4), o_is = take(N * 4), o_l = take(N * 4), o_r = take(N * 4), o_p = take(2 * N * 4 + 8);
    const size_t o_f = take(N * 4), o_la = take(N * 4), o_nl = take(N * 16), o_nh = take(N * 16), o_a = take(N * 4);
    const size_t o_lv = take(N * 4), o_c = take(N * 4);
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 64, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, st);
    const size_t o_tmp = take(sort_bytes > scan_bytes ? sort_bytes : scan_bytes);
    unsigned char *arena = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&arena, off, st);
    if (e != cudaSuccess) return e;
    p.st = (BdStats *)(arena + o_st);
    p.lo = (float4 *)(arena + o_lo), p.hi = (float4 *)(arena + o_hi);
    p.keys = (unsigned long long *)(arena + o_k), p.keys_sorted = (unsigned long long *)(arena + o_ks);
    p.idx = (uint32_t *)(arena + o_i), p.idx_sorted = (uint32_t *)(arena + o_is);
    p.left = (int *)(arena + o_l), p.right = (int *)(arena + o_r), p.parent = (int *)(arena + o_p);
    p.first = (uint32_t *)(arena + o_f), p.last = (uint32_t *)(arena + o_la);
    p.nlo = (float4 *)(arena + o_nl), p.nhi = (float4 *)(arena + o_nh), p.arrive = (unsigned int *)(arena + o_a);
    p.live = (uint32_t *)(arena + o_lv), p.compact = (uint32_t *)(arena + o_c);
    const unsigned B = 256, G = (unsigned)((n + B - 1) / B);
    bd_init<<<1, 1, 0, st>>>(p.st);
    bd_prim_boxes<<<G, B, 0, st>>>(p);
    bd_keys<<<G, B, 0, st>>>(p);
    size_t tb = sort_bytes;
    cub::DeviceRadixSort::SortPairs(arena + o_tmp, tb, p.keys, p.keys_sorted, p.idx, p.idx_sorted, (int)n, 0, 64, st);
    bd_karras<<<G, B, 0, st>>>(p);
    bd_fit<<<G, B, 0, st>>>(p);
    bd_live<<<G, B, 0, st>>>(p);
    tb = scan_bytes;
    cub::DeviceScan::ExclusiveSum(arena + o_tmp, tb, p.live, p.compact, (int)n, st);
    bd_emit<<<G, B, 0, st>>>(p);
    bd_top<<<1, 1, 0, st>>>(p);
    bd_slots<<<G, B, 0, st>>>(p);
    bd_depth<<<G, B, 0, st>>>(p);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, p.st, sizeof(BdStats), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(arena, st);
    return e;
}

#endif // RT_BVH_DEVICE_CUH
