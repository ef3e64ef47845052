// bvh_build.h — host-side build of the device BVH (replaces NewBVH, bvh.go:142-185).
#ifndef RT_BVH_BUILD_H
#define RT_BVH_BUILD_H

#include <cstdint>
#include <memory>
#include <new>
#include <utility>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_trace.h"

// Host copy of the scene's hittables (World.hittables split by kind, with their object IDs).
struct ScenePrims {
    std::vector<rt_sphere> spheres;
    std::vector<uint32_t> sphere_ids;
    std::vector<rt_quad> quads;
    std::vector<uint32_t> quad_ids;
    size_t size() const { return spheres.size() + quads.size(); }
};

// std::allocator that leaves trivially constructible elements uninitialised on resize(): the builder
// writes every element of these arrays itself, and zero-filling 70 MB first cost 18 ms of a 1 M-sphere build.
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U>
    struct rebind {
        typedef NoInitAlloc<U> other;
    };
    NoInitAlloc() = default;
    template <class U>
    NoInitAlloc(const NoInitAlloc<U> &) {}
    template <class U>
    void construct(U *p) { ::new ((void *)p) U; } // default-initialisation: no zero fill
    template <class U, class... A>
    void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
};
template <class T>
using RawVec = std::vector<T, NoInitAlloc<T>>;

struct FlatBvh {
    RawVec<F4> nodes;          // 2 x F4 per node (min.xyz, ref)(max.xyz, 0), siblings adjacent, depth-first order
    RawVec<F4> dev_nodes;      // the same nodes as the kernels read them: (centre.xyz, ref)(half-extent.xyz, 0),
                               // [c - h, c + h] encloses [min, max] (box_test in rt_trace.h); followed by the
                               // "walk pairs" of the leaf-start chains (below) when those were built
    // Leaf start (build_leaf_start): a ray that leaves a known primitive starts inside its leaf's box and inside
    // the boxes of all the leaf's ancestors, so those need no test — only the siblings along the path do.  Two
    // siblings are copied side by side into a "walk pair" (a pair of device nodes like any other) and the refs of a
    // leaf's walk pairs form its chain; the kernel pushes the chain, starts at the leaf and traverses as usual.
    RawVec<uint32_t> walk_src;   // 2 per walk pair: the real nodes copied into it (second = RT_REF_NONE: a never-hit dummy)
    RawVec<uint32_t> chains;     // per leaf: [len, len walk-pair refs (root side first), the leaf's own ref]
    RawVec<uint32_t> sph_chain;  // per sphere slot: offset of its leaf's chain in `chains`
    RawVec<uint32_t> quad_chain; // per quad slot: likewise
    RawVec<F4> sph;            // per sphere slot: centre, radius
    RawVec<I2> meta;           // per sphere slot: object ID, material index
    RawVec<uint32_t> sph_prim; // per sphere slot: index into ScenePrims.spheres
    RawVec<F4> quad;           // per quad slot: RT_QUAD_F4 x F4 (layout in rt_trace.h)
    RawVec<uint32_t> quad_prim; // per quad slot: index into ScenePrims.quads
    uint32_t root_ref = RT_REF_NONE;
    uint32_t max_depth = 0;         // deepest chain of inner nodes (bounds the traversal stack)
    float pad_min = 0, pad_max = 0; // smallest / largest box padding applied to a primitive
};

// max_leaf in [1, RT_MAX_LEAF].  origin_radius: rt_scene_desc.ray_origin_radius (0 = derive).
// `center` (optional): compute_scene_center's m, if the caller has it already.
void build_flat_bvh(const ScenePrims &prims, float origin_radius, int max_leaf, FlatBvh *out, const double *center = nullptr,
                    bool leaf_start = false);
// Recompute all boxes for a (larger) origin radius; topology and slot order are unchanged.
void refit_flat_bvh(const ScenePrims &prims, float origin_radius, FlatBvh *bvh);
// Per-axis median m of the primitive centres, the 90th percentile of |c - m| + extent, and
// (optional) max over primitives of |c - m| - extent: how far from m the nearest point of the farthest
// primitive is.
void compute_scene_center(const ScenePrims &prims, double m[3], double *extent90, double *surface_extent = nullptr);
// Fold each material's texture into its 32-byte device record (layout in rt_shade.h).
void pack_materials(const rt_scene_desc *d, std::vector<F4> *out);
// Copies (and validates the IDs of) the hittables of a scene description; false = bad IDs.
bool load_scene_prims(const rt_scene_desc *d, ScenePrims *out);

#endif
