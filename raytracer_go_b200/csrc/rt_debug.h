// rt_debug.h — bounds checks of the debug build (librt_b200_debug.so, csrc/Makefile `debug`).
//
// compute-sanitizer is closed on the GPU pool this library is developed on, so the debug build carries its own
// checks: every index the kernels form — traversal-stack depth, node / leaf-slot / chain references, hit slots,
// material, image, Perlin-table and texel indices, queue entries and radiance slots — is compared with the size of
// the array it addresses.  A violation is counted per check (g_dbg_violations), the access itself still happens; after
// the call has drained, the host turns a non-zero counter into RT_ERR_INTERNAL naming the check.  The release build
// compiles the checks away (RT_DBG expands to nothing).  RT_B200_DEBUG=1 makes the Python loader (lib.py) pick the
// debug library; `pytest -m gpu` then runs the whole parity suite under it (scripts/gpu_run.sh validate-debug).
#ifndef RT_DEBUG_H
#define RT_DEBUG_H

#include <stdint.h>

#ifndef RT_DEBUG_CHECKS
#define RT_DEBUG_CHECKS 0
#endif

enum {
    RT_DBG_STACK = 0,    // traversal stack deeper than the handle's stack_depth / RT_LOCAL_STACK
    RT_DBG_NODE = 1,     // inner reference outside the node array
    RT_DBG_LEAF = 2,     // leaf slots outside the sphere / quad slot arrays
    RT_DBG_CHAIN = 3,    // leaf-start chain outside the chain array
    RT_DBG_HIT_SLOT = 4, // hit slot outside the slot arrays when shading
    RT_DBG_MATERIAL = 5, // material index outside the material array
    RT_DBG_TEXTURE = 6,  // image / Perlin table index outside its array
    RT_DBG_TEXEL = 7,    // texel outside the image
    RT_DBG_QUEUE = 8,    // survivor-queue entry outside the queue
    RT_DBG_SAMPLE = 9,   // radiance slot outside the pass buffer
    RT_DBG_PIXEL = 10,   // pixel outside the image / accumulator
    RT_DBG_N = 12
};

struct DbgBounds {
    uint32_t n_nodes, n_slots, n_quad_slots, n_mats, n_chain_words, stack_entries, n_images, n_perlins;
    unsigned long long queue_cap, samples_cap, n_pixels;
};

#if RT_DEBUG_CHECKS && defined(__CUDACC__)
__device__ DbgBounds g_dbg_bounds;
__device__ unsigned int g_dbg_violations[RT_DBG_N];
#if defined(__CUDA_ARCH__)
#define RT_DBG(cond, code)                                      \
    do {                                                        \
        if (!(cond)) atomicAdd(&g_dbg_violations[(code)], 1u); \
    } while (0)
#define RT_DBG_B(field) (g_dbg_bounds.field)
#else
#define RT_DBG(cond, code) ((void)0)
#endif
#else
#define RT_DBG(cond, code) ((void)0)
#endif

static const char *const rt_dbg_names[RT_DBG_N] = {"traversal stack depth", "node reference", "leaf slot range",
                                                   "leaf-start chain",      "hit slot",       "material index",
                                                   "texture table index",   "texel index",    "queue entry",
                                                   "radiance slot",         "pixel index",    "(unused)"};

#endif // RT_DEBUG_H
