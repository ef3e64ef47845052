// rt_trace.h — closest-hit query over the flattened BVH with World.Hit semantics.
//
// Replaces BVH.Hit (bvh.go:220-249) + Aabb.Hit (bvh.go:52-102) + Sphere.Hit's root search
// (hittables.go:96-116).  The answer is defined by the reference's brute-force list
// (World.Hit, hittables.go:55-72): among all spheres, the smallest accepted root, the lowest
// object index winning exact ties.  That definition is independent of traversal order:
//   candidate(s) = near root if tmin < near root, else far root if tmin < far root  (strict,
//                  bvh.go:18-20), computed with the reference's float32 operation order;
//   result       = argmin over (candidate, object index), candidates >= tmax discarded.
// Box tests only cull; they are fused (fmaf) and conservative: boxes are padded on the host so
// no sphere whose float32 test accepts a root can be culled (see bvh_build.cpp).
//
// Device layout (uploaded once, depth-first order, siblings adjacent):
//   node i  = 2 x F4 = 32 bytes: (min.x, min.y, min.z, ref) (max.x, max.y, max.z, unused)
//   ref     = inner: index of the first of its two children (children are nodes ref, ref+1)
//             leaf : RT_LEAF | first_slot << 3 | (count-1)        (1..8 spheres), or
//                    RT_LEAF | RT_LEAF_QUAD | first_slot << 3 | (count-1)   (quads, own slot array)
//   sphere slot s = F4 (cx, cy, cz, r) + I2 (object ID, material index)
//   quad slot s   = 5 x F4: (Q.xyz, D) (u.xyz, bits material) (v.xyz, bits object ID) (w.xyz, 0)
//                   (normal.xyz, 0) — Q, u, v and the fields NewQuad derives (hittables.go:149-165)
#ifndef RT_TRACE_H
#define RT_TRACE_H

#include "rt_debug.h"
#include "rt_math.h"

#if defined(__CUDACC__)
#define RT_ALIGN(n) __align__(n)
#else
#define RT_ALIGN(n) alignas(n)
#endif

struct RT_ALIGN(16) F4 {
    float x, y, z, w;
};
struct RT_ALIGN(8) I2 {
    int32_t x, y;
};

#define RT_LEAF 0x80000000u
#define RT_LEAF_QUAD 0x40000000u      /* with RT_LEAF: the leaf holds quads */
#define RT_LEAF_SLOT_MASK 0x3FFFFFFFu /* clears the two flag bits */
#define RT_REF_NONE 0xFFFFFFFFu       /* empty scene / stack bottom */
#define RT_MAX_LEAF 8
#define RT_QUAD_F4 5                  /* F4 per quad slot */
#define RT_HIT_QUAD 0x40000000u       /* HitRec.slot flag: the slot indexes the quad array */

struct HitRec {
    float t;
    uint32_t slot; // RT_REF_NONE on miss; RT_HIT_QUAD | quad slot for a quad
};

struct WorkCounters {
    unsigned long long box_tests, sphere_tests;
};

// Per-thread traversal stack in thread-private (local) memory.
template <int N>
struct LocalStack {
    uint32_t e[N];
    int sp;
    RT_HD void reset() { sp = 0; }
    RT_HD void push(uint32_t r) {
        RT_DBG(sp < N, RT_DBG_STACK);
        e[sp++] = r;
    }
    RT_HD uint32_t pop() { return sp > 0 ? e[--sp] : RT_REF_NONE; }
    RT_HD void push_if(bool c, uint32_t r) {
        if (c) push(r);
    }
    // `keep` when have_next, else the popped entry (RT_REF_NONE when the stack is empty)
    RT_HD uint32_t next_or_pop(bool have_next, uint32_t keep) { return have_next ? keep : pop(); }
};

// Per-thread traversal stack in shared memory: entry d of thread t lives at base[d*stride + t],
// so a warp's accesses at one depth hit 32 distinct banks.
struct StridedStack {
    uint32_t *base; // already offset by the thread index
    int stride;
#if defined(__CUDA_ARCH__)
    // 32-bit shared-window addresses: push / pop are one STS / LDS plus one add (a generic pointer
    // makes ptxas carry the generic and the shared address side by side).  The asm statements are
    // volatile, so they keep their order; nothing else touches the stack memory.
    uint32_t base_s, top_s;
    __device__ __forceinline__ void reset() { base_s = top_s = (uint32_t)__cvta_generic_to_shared(base); }
    __device__ __forceinline__ void push(uint32_t r) {
        RT_DBG(top_s - base_s < 4u * (uint32_t)stride * RT_DBG_B(stack_entries), RT_DBG_STACK);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(top_s), "r"(r));
        top_s += 4u * (uint32_t)stride;
    }
    __device__ __forceinline__ uint32_t pop() {
        if (top_s == base_s) return RT_REF_NONE;
        top_s -= 4u * (uint32_t)stride;
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(top_s));
        return r;
    }
    // Predicated forms for the inner traversal step: the lanes of an incoherent warp disagree on
    // "both children hit / one / none" almost every step, so branches there only add BSSY / BRA /
    // BSYNC around code that is executed anyway (control flow was 15 % of the issued instructions).
    __device__ __forceinline__ void push_if(bool c, uint32_t r) {
        RT_DBG(!c || top_s - base_s < 4u * (uint32_t)stride * RT_DBG_B(stack_entries), RT_DBG_STACK);
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u32 [%0], %1; }" ::"r"(top_s), "r"(r), "r"((uint32_t)c));
        top_s += c ? 4u * (uint32_t)stride : 0u;
    }
    __device__ __forceinline__ uint32_t next_or_pop(bool have_next, uint32_t keep) {
        const bool do_pop = !have_next && top_s != base_s;
        top_s -= do_pop ? 4u * (uint32_t)stride : 0u;
        uint32_t r = have_next ? keep : RT_REF_NONE;
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p ld.shared.u32 %0, [%1]; }" : "+r"(r) : "r"(top_s), "r"((uint32_t)do_pop));
        return r;
    }
#else
    int sp;
    void reset() { sp = 0; }
    void push(uint32_t r) { base[sp++ * stride] = r; }
    uint32_t pop() { return sp == 0 ? RT_REF_NONE : base[--sp * stride]; }
    void push_if(bool c, uint32_t r) {
        if (c) push(r);
    }
    uint32_t next_or_pop(bool have_next, uint32_t keep) { return have_next ? keep : pop(); }
#endif
};

RT_HD float rt_fmin(float a, float b) { return fminf(a, b); } // NaN-ignoring: a NaN slab never culls
RT_HD float rt_fmax(float a, float b) { return fmaxf(a, b); }
// Three-input min/max: one FMNMX3 on sm_100 (PTX max.f32 d,a,b,c); same NaN-ignoring semantics.
RT_HD float rt_fmin3(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
#else
    return fminf(fminf(a, b), c);
#endif
}
RT_HD float rt_fmax3(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}

// Reciprocal of a direction component for culling only.  A zero / denormal component would give
// inf and then inf - inf = NaN inside the fused slab test, so it is replaced by +-1e30: a ray
// parallel to a slab then sees (-huge, +huge) when its origin is inside the slab and an empty
// interval when outside, which is the geometric answer.
RT_HD float cull_rcp(float x) {
    if (fabsf(x) < 1e-30f) return copysignf(1e30f, x);
#if defined(__CUDA_ARCH__)
    return __fdividef(1.0f, x);
#else
    return 1.0f / x;
#endif
}

// Slab test of one child box against (tmin, tbest]; returns the entry distance in tnear.
// The box is stored as centre c and half-extent h (device_nodes() in bvh_build.cpp), so the near and
// far planes of an axis are m -+ h*|inv| with m = (c - o)*inv: no per-axis min/max to order them.
// Per box 9 FFMA + 4 FMNMX(3) + 2 FSETP, against 6 FFMA + 10 FMNMX(3) + FSETP for a (min, max)
// box: the min/max instructions run on the half-rate ALU pipe, which is the busiest pipe of the
// traversal loop (ncu: alu 62 %, fma 24 % of peak), the FFMAs on the FMA pipe.
// inv = 1/d, noi = -(o * inv), ainv = |inv|.  Conservative, not the reference's arithmetic.
RT_HD bool box_test(const F4 &c, const F4 &h, V3 inv, V3 noi, V3 ainv, float tmin, float tbest, float &tnear) {
    const float mx = fmaf(c.x, inv.x, noi.x), my = fmaf(c.y, inv.y, noi.y), mz = fmaf(c.z, inv.z, noi.z);
    const float nx = fmaf(-h.x, ainv.x, mx), ny = fmaf(-h.y, ainv.y, my), nz = fmaf(-h.z, ainv.z, mz);
    const float fx = fmaf(h.x, ainv.x, mx), fy = fmaf(h.y, ainv.y, my), fz = fmaf(h.z, ainv.z, mz);
    const float tn = rt_fmax(rt_fmax3(nx, ny, nz), tmin);
    const float tf = rt_fmin(rt_fmin3(fx, fy, fz), tbest);
    tnear = tn;
    return tn <= tf;
}

#if defined(__CUDA_ARCH__)
// The same slab test for BOTH children of a pair with sm_100's packed FP32 instructions (FFMA2: two fused
// multiply-adds per lane and issue slot).  The shared-memory copy of a pair is stored transposed for it
// (stage_scene in rt_kernels.cuh):
//   w0 = (cL.x, cL.y, cR.x, cR.y)   w1 = (cL.z, cR.z, hL.z, hR.z)   w2 = (hL.x, hL.y, hR.x, hR.y)   w3 = (refL, refR, -, -)
// so the x/y components of a box pair with the ray's (inv.x, inv.y) and the two z components with (inv.z, inv.z):
// 9 FFMA2 instead of 18 FFMA per pair, the same fused operation per component, hence bit-identical results.
// The kernels are bound by instruction issue, not by the FMA pipe (ncu: fma 33 %, issue 80 %).
struct RayPairs {
    float2 inv_xy, inv_zz, noi_xy, noi_zz, ainv_xy, ainv_zz;
};
__device__ __forceinline__ void box_test_pair(const F4 &w0, const F4 &w1, const F4 &w2, const RayPairs &r, float tmin,
                                              float tbest, bool &hl, bool &hr, float &tl, float &tr) {
    const float2 mL = __ffma2_rn(make_float2(w0.x, w0.y), r.inv_xy, r.noi_xy);
    const float2 mR = __ffma2_rn(make_float2(w0.z, w0.w), r.inv_xy, r.noi_xy);
    const float2 mZ = __ffma2_rn(make_float2(w1.x, w1.y), r.inv_zz, r.noi_zz);
    const float2 nL = __ffma2_rn(make_float2(-w2.x, -w2.y), r.ainv_xy, mL), fL = __ffma2_rn(make_float2(w2.x, w2.y), r.ainv_xy, mL);
    const float2 nR = __ffma2_rn(make_float2(-w2.z, -w2.w), r.ainv_xy, mR), fR = __ffma2_rn(make_float2(w2.z, w2.w), r.ainv_xy, mR);
    const float2 nZ = __ffma2_rn(make_float2(-w1.z, -w1.w), r.ainv_zz, mZ), fZ = __ffma2_rn(make_float2(w1.z, w1.w), r.ainv_zz, mZ);
    tl = rt_fmax(rt_fmax3(nL.x, nL.y, nZ.x), tmin);
    tr = rt_fmax(rt_fmax3(nR.x, nR.y, nZ.y), tmin);
    hl = tl <= rt_fmin(rt_fmin3(fL.x, fL.y, fZ.x), tbest);
    hr = tr <= rt_fmin(rt_fmin3(fR.x, fR.y, fZ.y), tbest);
}
#endif

// hittables.go:96-116 for one sphere, in the reference's operation order (unfused).
// a = |d|^2 (hittables.go:98) is hoisted: it does not depend on the sphere.
RT_HD bool sphere_candidate(const F4 &s, V3 o, V3 d, float a, float tmin, float &t_out) {
    V3 oc = o - v3(s.x, s.y, s.z);            // hittables.go:97
    float half_b = dot(d, oc);                // :99
    float c = lensq(oc) - s.w * s.w;          // :100
    float disc = half_b * half_b - a * c;     // :102
    if (disc < 0) return false;               // :104
    float sqt = sqrt32(disc);                 // :108
    float root = div32(-half_b - sqt, a);     // :110
    if (!(tmin < root)) {
        root = div32(-half_b + sqt, a);       // :112
        if (!(tmin < root)) return false;
    }
    t_out = root;
    return true;
}

// hittables.go:167-190 for one quad, in the reference's operation order.  A root beyond tbest is
// dropped before the in-plane test (the reference tests the interval first too, :176).
RT_HD bool quad_candidate(const F4 *__restrict__ q, V3 o, V3 d, float tmin, float tbest, float &t_out) {
    const V3 n = v3(q[4].x, q[4].y, q[4].z);
    const float denom = dot(d, n);                            // :168
    if (fabs((double)denom) < 1e-8) return false;             // :170
    const float t = div32(q[0].w - dot(n, o), denom);         // :174
    if (!(tmin < t) || !(t <= tbest)) return false;           // :176 (ties are resolved by the caller)
    const V3 p = d * t + o;                                   // :180
    const V3 ph = p - v3(q[0].x, q[0].y, q[0].z);             // :181
    const V3 w = v3(q[3].x, q[3].y, q[3].z);
    const float alpha = dot(w, cross(ph, v3(q[2].x, q[2].y, q[2].z))); // :182
    const float beta = dot(w, cross(v3(q[1].x, q[1].y, q[1].z), ph));  // :183
    if (alpha < 0 || 1 < alpha || beta < 0 || 1 < beta) return false;  // :185, 192-194
    t_out = t;
    return true;
}

// Object ID of a hit slot (only needed to break exact ties).
RT_HD int32_t slot_object_id(uint32_t slot, const I2 *__restrict__ meta, const F4 *__restrict__ quads) {
    if (slot & RT_HIT_QUAD) return (int32_t)as_uint(quads[(size_t)RT_QUAD_F4 * (slot & ~RT_HIT_QUAD) + 2].w);
    return meta[slot].x;
}

// PRED: the inner step (both children hit / one / none) as straight-line predicated code instead
// of branches.  Pays for incoherent warps, whose lanes disagree nearly every step (secondary
// megakernel, shared-memory stack: +1.4 % on C2); costs for coherent warps and for the
// local-memory stack (C4: -3.5 %), which keep the branches.
//
// Leaf start (chains != nullptr and start != RT_REF_NONE): the ray leaves a primitive whose leaf is known.  The
// leaf's box and the boxes of all its ancestors contain the origin, so the slab test accepts them for any
// direction: they are skipped.  What remains to be tested are the siblings along the path; the host has copied
// them, two by two, into "walk pairs" that look like any other pair of nodes (bvh_build.h).  The chain of the
// leaf is pushed (root side first), traversal starts at the leaf itself and then pops the walk pairs from the
// deepest up — the order the top-down traversal would have visited them in.  Skipping a box test can only add
// candidates, never remove one, so the closest hit is unchanged (and so is every bit of the image).
// NSTRIDE: bytes from one node to the next in `nodes`.  32 in global memory.  The shared-memory copy uses 40: a
// pair of nodes then starts every 80 bytes, so the first 16-byte word of pair p lies in bank group 5p mod 8 —
// all eight groups — instead of 4p mod 8 (two groups): the quarter-warp of an LDS.128 of 8 random nodes collides
// far less (ncu r02e: 38 % of the shared wavefronts were conflict replays with the dense layout).
// PACKED: `nodes` is the transposed shared-memory copy (box_test_pair above); device code only.
template <class Stack, bool COUNT, bool QUADS = false, bool PRED = false, int NSTRIDE = 32, bool PACKED = false>
RT_HD void trace_closest(const F4 *__restrict__ nodes, const F4 *__restrict__ sph,
                         const I2 *__restrict__ meta, uint32_t root_ref, V3 o, V3 d, float tmin,
                         float tmax, Stack &stack, HitRec &hit, WorkCounters *wc,
                         const F4 *__restrict__ quads = nullptr, const uint32_t *__restrict__ chains = nullptr,
                         uint32_t start = RT_REF_NONE) {
    const V3 inv = v3(cull_rcp(d.x), cull_rcp(d.y), cull_rcp(d.z));
    const V3 noi = v3(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
    const V3 ainv = v3(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
    const float a = lensq(d);
#if defined(__CUDA_ARCH__)
    RayPairs rp;
    if constexpr (PACKED) {
        rp.inv_xy = make_float2(inv.x, inv.y), rp.inv_zz = make_float2(inv.z, inv.z);
        rp.noi_xy = make_float2(noi.x, noi.y), rp.noi_zz = make_float2(noi.z, noi.z);
        rp.ainv_xy = make_float2(ainv.x, ainv.y), rp.ainv_zz = make_float2(ainv.z, ainv.z);
    }
#endif
    float tbest = tmax;
    uint32_t best_slot = RT_REF_NONE;
    int32_t best_id = 0x7fffffff;
    bool have_id = false; // best_id is loaded lazily: only exact ties need it
    stack.reset();
    uint32_t ref = root_ref;
    if (chains != nullptr && start != RT_REF_NONE) {
        RT_DBG(start < RT_DBG_B(n_chain_words), RT_DBG_CHAIN);
        const uint32_t len = chains[start];
        RT_DBG(start + len + 1 < RT_DBG_B(n_chain_words), RT_DBG_CHAIN);
        for (uint32_t k = 1; k <= len; k++) stack.push(chains[start + k]);
        ref = chains[start + len + 1];
    }
    for (;;) {
        // "while-while": every lane first descends through inner nodes until it holds a leaf (or
        // nothing), then the lanes test their leaves together.  RT_REF_NONE has the leaf bit set.
        while (!(ref & RT_LEAF)) {
            RT_DBG(!(ref & 1u) && ref + 1 < RT_DBG_B(n_nodes), RT_DBG_NODE);
            const F4 *np = reinterpret_cast<const F4 *>(reinterpret_cast<const char *>(nodes) + (size_t)ref * NSTRIDE);
            float tl, tr;
            bool hl, hr;
            uint32_t lref, rref;
#if defined(__CUDA_ARCH__)
            if constexpr (PACKED) {
                const F4 w0 = np[0], w1 = np[1], w2 = np[2];
                const uint2 refs = *reinterpret_cast<const uint2 *>(np + 3);
                box_test_pair(w0, w1, w2, rp, tmin, tbest, hl, hr, tl, tr);
                lref = refs.x, rref = refs.y;
            } else
#endif
            {
                const F4 l0 = np[0], l1 = np[1];
                const F4 r0 = np[2], r1 = np[3]; // (the sibling follows its node directly for either stride)
                hl = box_test(l0, l1, inv, noi, ainv, tmin, tbest, tl);
                hr = box_test(r0, r1, inv, noi, ainv, tmin, tbest, tr);
                lref = as_uint(l0.w), rref = as_uint(r0.w);
            }
            if (COUNT) wc->box_tests += 2;
            // nearer child first, the other one (if hit) on the stack; nothing hit: pop
            if (PRED) {
                const bool take_l = hl && (tl <= tr || !hr);
                stack.push_if(hl && hr, take_l ? rref : lref);
                ref = stack.next_or_pop(hl || hr, take_l ? lref : rref);
            } else if (hl && hr) {
                const bool left_first = tl <= tr;
                stack.push(left_first ? rref : lref);
                ref = left_first ? lref : rref;
            } else if (hl) {
                ref = lref;
            } else if (hr) {
                ref = rref;
            } else {
                ref = stack.pop();
            }
        }
        if (ref == RT_REF_NONE) break;
        const uint32_t first = (ref & RT_LEAF_SLOT_MASK) >> 3, count = (ref & 7u) + 1;
        RT_DBG(first + count <= ((ref & RT_LEAF_QUAD) ? RT_DBG_B(n_quad_slots) : RT_DBG_B(n_slots)), RT_DBG_LEAF);
        if (QUADS && (ref & RT_LEAF_QUAD)) {
            for (uint32_t s = first; s < first + count; s++) {
                float t;
                if (COUNT) wc->sphere_tests += 1;
                if (!quad_candidate(quads + (size_t)RT_QUAD_F4 * s, o, d, tmin, tbest, t)) continue;
                const uint32_t hs = s | RT_HIT_QUAD;
                if (t < tbest) {
                    tbest = t, best_slot = hs, have_id = false;
                } else if (t == tbest && best_slot != RT_REF_NONE) {
                    if (!have_id) best_id = slot_object_id(best_slot, meta, quads), have_id = true;
                    const int32_t id = slot_object_id(hs, meta, quads);
                    if (id < best_id) best_slot = hs, best_id = id;
                }
            }
        } else {
            for (uint32_t s = first; s < first + count; s++) {
                const F4 sp = sph[s];
                float t;
                if (COUNT) wc->sphere_tests += 1;
                if (!sphere_candidate(sp, o, d, a, tmin, t)) continue;
                if (t < tbest) {
                    tbest = t, best_slot = s, have_id = false;
                } else if (t == tbest && best_slot != RT_REF_NONE) {
                    // exact tie: World.Hit keeps the earlier object (strict `<`, hittables.go:59-69)
                    if (!have_id) best_id = QUADS ? slot_object_id(best_slot, meta, quads) : meta[best_slot].x, have_id = true;
                    const int32_t id = meta[s].x;
                    if (id < best_id) best_slot = s, best_id = id;
                }
            }
        }
        ref = stack.pop();
    }
    hit.t = tbest;
    hit.slot = best_slot;
}

// Closest hit among an explicit list of candidate slots (bit 30 = quad slot, as HitRec.slot): the leaf loop of
// trace_closest without the tree.  World.Hit's answer is the argmin over (accepted root, object index) of ALL
// primitives, so it is also the argmin over any SUPERSET of the primitives the ray can hit — which is what the
// per-pixel candidate lists of the primary stage are (pixel_candidates_kernel in rt_kernels.cuh).
template <bool COUNT, bool QUADS>
RT_HD void trace_candidates(const uint32_t *__restrict__ list, uint32_t n, const F4 *__restrict__ sph,
                            const I2 *__restrict__ meta, const F4 *__restrict__ quads, V3 o, V3 d, float tmin, float tmax,
                            HitRec &hit, WorkCounters *wc) {
    const float a = lensq(d);
    float tbest = tmax;
    uint32_t best_slot = RT_REF_NONE;
    int32_t best_id = 0x7fffffff;
    bool have_id = false;
    if (COUNT) wc->sphere_tests += n;
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t hs = list[k];
        RT_DBG((hs & ~RT_HIT_QUAD) < ((hs & RT_HIT_QUAD) ? RT_DBG_B(n_quad_slots) : RT_DBG_B(n_slots)), RT_DBG_LEAF);
        float t;
        if (QUADS && (hs & RT_HIT_QUAD)) {
            if (!quad_candidate(quads + (size_t)RT_QUAD_F4 * (hs & ~RT_HIT_QUAD), o, d, tmin, tbest, t)) continue;
        } else {
            const F4 sp = sph[hs];
            if (!sphere_candidate(sp, o, d, a, tmin, t)) continue;
        }
        if (t < tbest) {
            tbest = t, best_slot = hs, have_id = false;
        } else if (t == tbest && best_slot != RT_REF_NONE) { // exact tie: the earlier object (hittables.go:59-69)
            if (!have_id) best_id = slot_object_id(best_slot, meta, quads), have_id = true;
            const int32_t id = slot_object_id(hs, meta, quads);
            if (id < best_id) best_slot = hs, best_id = id;
        }
    }
    hit.t = tbest;
    hit.slot = best_slot;
}

// ---------------------------------------------------------------------------------------------
// beams: every camera ray of one pixel at once (interval arithmetic), for the per-pixel candidate lists
// ---------------------------------------------------------------------------------------------
// All rays of a pixel start in the box [olo, ohi] (the defocus disk around the camera centre) and have directions in
// the box [dlo, dhi] (pixel footprint minus origin box).  At ray parameter t >= 0 every one of them is inside the box
// [olo + t*dlo, ohi + t*dhi]; a ray that passes the kernels' slab test of a node box is inside that node box at some t,
// so the two boxes overlap on all three axes at that t.  beam_box_test asks whether such a t exists: per axis two
// half-lines in t (olo + t*dlo <= max and ohi + t*dhi >= min), intersected.  The node box is widened by RT_BEAM_EPS
// of the magnitudes involved — two orders of magnitude more than the rounding of box_test (a few 2^-23 of
// |c*inv| + |o*inv| + h*|inv|, approximate reciprocal included).  A box the beam test rejects is rejected by box_test
// for every ray of the pixel, so the subtree below it cannot contribute a hit.
#define RT_BEAM_EPS 1e-5f
struct Beam {
    V3 olo, ohi, dlo, dhi;
};
// one axis: narrows [t0, t1]; false = the beam never overlaps the slab
RT_HD bool beam_axis(float c, float h, float olo, float ohi, float dlo, float dhi, float &t0, float &t1) {
    const float e = RT_BEAM_EPS * (fabsf(c) + h + fmaxf(fabsf(olo), fabsf(ohi))) + 1e-30f;
    const float A = (c + h + e) - olo; // olo + t*dlo <= max
    const float B = (c - h - e) - ohi; // ohi + t*dhi >= min
    if (dlo > 0.0f) t1 = fminf(t1, A / dlo);
    else if (dlo < 0.0f) t0 = fmaxf(t0, A / dlo);
    else if (A < 0.0f) return false;
    if (dhi > 0.0f) t0 = fmaxf(t0, B / dhi);
    else if (dhi < 0.0f) t1 = fminf(t1, B / dhi);
    else if (B > 0.0f) return false;
    return true;
}
RT_HD bool beam_box_test(const F4 &c, const F4 &h, const Beam &b) {
    float t0 = 0.0f, t1 = INFINITY;
    if (!beam_axis(c.x, h.x, b.olo.x, b.ohi.x, b.dlo.x, b.dhi.x, t0, t1)) return false;
    if (!beam_axis(c.y, h.y, b.olo.y, b.ohi.y, b.dlo.y, b.dhi.y, t0, t1)) return false;
    if (!beam_axis(c.z, h.z, b.olo.z, b.ohi.z, b.dlo.z, b.dhi.z, t0, t1)) return false;
    return !(t0 > t1 * (1.0f + RT_BEAM_EPS)); // (NaN never culls: fminf / fmaxf ignore it)
}

// Walks the tree (global-memory layout, 32 bytes per node) with a beam and writes the slots of every leaf it reaches
// (bit 30 set for quad slots) to out[0..cap); returns how many there are — more than cap means the list overflowed.
template <bool QUADS>
RT_HD uint32_t beam_candidates(const F4 *__restrict__ nodes, uint32_t root_ref, const Beam &b, uint32_t *out, uint32_t cap) {
    uint32_t n = 0;
    LocalStack<64> stack;
    stack.reset();
    uint32_t ref = root_ref;
    for (;;) {
        while (!(ref & RT_LEAF)) {
            const F4 *np = nodes + 2 * (size_t)ref;
            const F4 l0 = np[0], l1 = np[1], r0 = np[2], r1 = np[3];
            const bool hl = beam_box_test(l0, l1, b), hr = beam_box_test(r0, r1, b);
            const uint32_t lref = as_uint(l0.w), rref = as_uint(r0.w);
            if (hl && hr) {
                stack.push(rref);
                ref = lref;
            } else if (hl) {
                ref = lref;
            } else if (hr) {
                ref = rref;
            } else {
                ref = stack.pop();
            }
        }
        if (ref == RT_REF_NONE) break;
        const uint32_t first = (ref & RT_LEAF_SLOT_MASK) >> 3, count = (ref & 7u) + 1;
        const uint32_t flag = (QUADS && (ref & RT_LEAF_QUAD)) ? RT_HIT_QUAD : 0u;
        for (uint32_t s = first; s < first + count; s++) {
            if (n < cap) out[n] = s | flag;
            n++;
        }
        if (n > cap) return n; // overflow: the caller falls back to the tree, no point in walking on (a beam along the
                               // horizon of a million-sphere plane would otherwise visit thousands of leaves)
        ref = stack.pop();
    }
    return n;
}

#endif // RT_TRACE_H
