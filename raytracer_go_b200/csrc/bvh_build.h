// bvh_build.h — host-side build of the device BVH (replaces NewBVH, bvh.go:142-185).
#ifndef RT_BVH_BUILD_H
#define RT_BVH_BUILD_H

#include <cstdint>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_trace.h"

struct FlatBvh {
    std::vector<F4> nodes;          // 2 x F4 per node, siblings adjacent, depth-first order
    std::vector<F4> sph;            // per slot: centre, radius
    std::vector<I2> meta;           // per slot: object index, material index
    uint32_t root_ref = RT_REF_NONE;
    uint32_t max_depth = 0;         // deepest chain of inner nodes (bounds the traversal stack)
    float pad_min = 0, pad_max = 0; // smallest / largest box padding applied to a sphere
};

// max_leaf in [1, RT_MAX_LEAF].  origin_radius: rt_scene_desc.ray_origin_radius (0 = derive).
void build_flat_bvh(const rt_sphere *spheres, uint64_t n, float origin_radius, int max_leaf, FlatBvh *out);
// Recompute all boxes for a (larger) origin radius; topology and slot order are unchanged.
void refit_flat_bvh(const rt_sphere *spheres, uint64_t n, float origin_radius, FlatBvh *bvh);
// Per-axis median of the sphere centres and the 90th percentile of |c - m| + r.
void compute_scene_center(const rt_sphere *spheres, uint64_t n, double m[3], double *extent90);

// Fold each material's texture into its 32-byte device record (layout in rt_shade.h).
void pack_materials(const rt_scene_desc *d, std::vector<F4> *out);

#endif
