// rt_b200.cu — the C ABI of librt_b200.so (include/rt_b200.h).
//
// Host side of the library: error reporting, the per-device workspace, the scene handle (host BVH
// build, flattening, upload), kernel launches and every extern "C" entry point.  The kernels are in
// rt_kernels.cuh; the per-ray arithmetic in rt_trace.h / rt_shade.h / rt_rng.h / rt_math.h.
//
// Compiled with -fmad=false: see rt_math.h.  No tensor cores (not a contraction), no RT cores
// (B200 has none).  There is no CPU path: every entry point that computes needs an sm_100 device.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <system_error>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/rt_b200.h"
#include "bvh_build.h"
#include "rt_math.h"
#include "rt_rng.h"
#include "rt_shade.h"
#include "rt_trace.h"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY : RT_ERR_CUDA,     \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define RC(call)                 \
    do {                         \
        int rc__ = (call);       \
        if (rc__ != RT_OK) return rc__; \
    } while (0)

#include "rt_kernels.cuh"
#include "bvh_device.cuh"

// ---------------------------------------------------------------------------------------------
// host side: per-device workspace, shared by all scene handles of the process
// ---------------------------------------------------------------------------------------------
// The per-pass radiance buffer (up to 1 GiB), the accumulator, the RGB8 image and its pinned
// staging copy are cached per device for the life of the process (rt_workspace_release frees
// them): a Camera.Render-style call creates and destroys a scene handle every time, and paying a
// half-gigabyte cudaMalloc/cudaFree per call costs more than the render itself.  The mutex is held
// for the duration of a render or resolve, which serialises calls on one device (they would
// serialise on the GPU anyway) and keeps distinct handles independent as far as results go.
#define RT_MAX_DEVICES 64
struct Workspace {
    std::mutex mu;
    float4 *samples = nullptr;
    size_t samples_cap = 0; // elements
    float *accum = nullptr;
    size_t accum_cap = 0; // floats
    uint8_t *rgb = nullptr, *h_rgb = nullptr;
    size_t rgb_cap = 0, h_rgb_cap = 0; // bytes
    float *h_accum = nullptr;
    size_t h_accum_cap = 0; // floats (pinned)
    float4 *queue = nullptr; // staged mode: 2 (ping-pong) x 3 x capacity float4 (origin, direction, throughput)
    size_t queue_cap = 0;    // elements in total
    uint32_t *lists = nullptr; // per-pixel candidate lists of the primary stage (RT_LIST_WORDS words per pixel of a pixel tile)
    size_t lists_cap = 0;      // words
    unsigned char *gather = nullptr; // rt_render_multi: tile-split image (+ sums) assembled on the root / sample-split receive buffer
    size_t gather_cap = 0;           // bytes
};
static Workspace g_ws[RT_MAX_DEVICES];

template <class T>
static int ws_reserve(T *&ptr, size_t &cap, size_t need, bool pinned_host = false) {
    if (need <= cap) return RT_OK;
    if (ptr) {
        if (pinned_host) cudaFreeHost(ptr);
        else cudaFree(ptr);
    }
    ptr = nullptr, cap = 0;
    if (pinned_host) CU(cudaMallocHost((void **)&ptr, need * sizeof(T)));
    else CU(cudaMalloc((void **)&ptr, need * sizeof(T)));
    cap = need;
    return RT_OK;
}

extern "C" void rt_workspace_release(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    for (int d = 0; d < n && d < RT_MAX_DEVICES; d++) {
        if (device >= 0 && d != device) continue;
        Workspace &w = g_ws[d];
        std::lock_guard<std::mutex> lock(w.mu);
        cudaSetDevice(d);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        cudaGetLastError();
        if (!w.samples && !w.accum && !w.rgb && !w.h_rgb && !w.h_accum && !w.queue && !w.gather && !w.lists) continue;
        cudaFree(w.samples), cudaFree(w.accum), cudaFree(w.rgb), cudaFree(w.queue), cudaFree(w.gather), cudaFree(w.lists);
        w.queue = nullptr, w.queue_cap = 0, w.gather = nullptr, w.gather_cap = 0, w.lists = nullptr, w.lists_cap = 0;
        if (w.h_rgb) cudaFreeHost(w.h_rgb);
        if (w.h_accum) cudaFreeHost(w.h_accum);
        w.samples = nullptr, w.accum = nullptr, w.rgb = nullptr, w.h_rgb = nullptr, w.h_accum = nullptr;
        w.samples_cap = w.accum_cap = w.rgb_cap = w.h_rgb_cap = w.h_accum_cap = 0;
    }
}

// Scene buffers come from the device's stream-ordered memory pool, configured to keep freed
// memory (release threshold = max): a Camera.Render-style caller creates and destroys a scene per
// frame, and cudaFree of a 17 MB texture was measured at 100-700 ms next to a multi-GiB workspace
// (profiles/r01ah_e2e_breakdown.txt); from the pool, create + destroy cost microseconds after the
// first frame.  rt_workspace_release() trims the pool.
static std::once_flag g_pool_once[RT_MAX_DEVICES];
static int scene_alloc(void **p, size_t bytes, int device, cudaStream_t st) {
    std::call_once(g_pool_once[device], [device]() {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    });
    CU(cudaMallocAsync(p, bytes, st));
    return RT_OK;
}
template <class T>
static int scene_alloc(T **p, size_t bytes, int device, cudaStream_t st) {
    return scene_alloc((void **)p, bytes, device, st);
}
static void scene_free(void *p, cudaStream_t st) {
    if (p) cudaFreeAsync(p, st);
}

// ---------------------------------------------------------------------------------------------
// host side: scene handle
// ---------------------------------------------------------------------------------------------
struct rt_scene {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    ScenePrims prims;
    FlatBvh bvh;
    double center[3] = {0, 0, 0};
    double extent90 = 0;
    double surface_extent = 0; // max over primitives of (distance of its nearest point from the centre)
    float origin_radius = 0;   // rays are exact when they start within this distance of the surface they hit
    bool fixed_radius = false; // the caller gave ray_origin_radius: an envelope, never enlarged
    DevScene dev{};
    bool use_smem = false;
    int block = 512;
    int pblock = 512;
    int minb = 2;
    bool use_split = false; // two-stage mode: coherent primary stage + megakernel on the survivors
    unsigned int *d_queue_count = nullptr; // RT_MAX_STAGES counters, one per stage queue
    int n_stages = 1;                      // coherent stages before the megakernel (RT_B200_STAGES)
    int primary_grid[2][2] = {{0, 0}, {0, 0}}; // [COUNT][FIRST]
    size_t primary_smem[2][2] = {{0, 0}, {0, 0}};
    // device buffers
    F4 *d_nodes = nullptr, *d_sph = nullptr, *d_mats = nullptr, *d_quads = nullptr;
    uint32_t *d_chains = nullptr; // leaf-start chains + per-slot chain offsets (one buffer)
    bool has_quads = false;
    I2 *d_meta = nullptr;
    DevImage *d_images = nullptr;
    DevPerlin *d_perlins = nullptr;
    std::vector<uint16_t *> d_texels;
    uint32_t n_images = 0, n_perlins = 0;
    size_t n_prims = 0;            // hittables of the scene (the host copy `prims` is empty for a device-built tree)
    bool device_built = false;     // the BVH was built on the GPU (bvh_device.cuh); `bvh` holds no host arrays then
    rt_sphere *d_raw = nullptr;    // device-built: the caller's spheres and IDs, kept for rebuilds with a larger radius
    uint32_t *d_raw_ids = nullptr;
    BdStats dev_stats{};           // device-built: statistics of the last build
    int grid_cache[2] = {0, 0};  // persistent grid size of the plain / counting megakernel
    size_t smem_cache[2] = {0, 0};
    unsigned int *d_counter = nullptr;
    unsigned long long *d_stats = nullptr;
    cudaStream_t stream = nullptr;     // stream in use
    cudaStream_t own_stream = nullptr; // created with the handle
    std::vector<cudaEvent_t> events;   // pairs around megakernel launches
};

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// No exception crosses the C ABI: host allocations (scene copies, BVH build, staging buffers) and thread
// creation can throw; the caller gets a status code and rt_last_error() instead.
template <class F>
static int guarded(F f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return fail(RT_ERR_OUT_OF_MEMORY, "host memory allocation failed");
    } catch (const std::exception &e) {
        return fail(RT_ERR_INTERNAL, "internal error: %s", e.what());
    } catch (...) {
        return fail(RT_ERR_INTERNAL, "internal error");
    }
}

extern "C" const char *rt_last_error(void) { return g_err.c_str(); }
extern "C" int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

extern "C" int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

static int select_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(RT_ERR_NO_DEVICE, "no CUDA device visible; librt_b200 has no CPU path");
    }
    if (device < 0 || device >= n) return fail(RT_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device, n);
    int major = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return fail(RT_ERR_NO_DEVICE, "device %d is sm_%dx; this library is built for sm_100a only", device, major);
    CU(cudaSetDevice(device));
    return RT_OK;
}

// ok(i) for every i in [0, n): RT_OK, or the status (and message) of the FIRST element that fails.  Arrays of a million
// elements (one material and texture per sphere in the stress scene) are checked on up to 8 threads — 7 ms of a 119 ms
// C4 frame otherwise; the workers only find the index, the caller's thread repeats that one check for the message.
template <class F>
static int check_all(uint64_t n, F ok) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned parts = n >= (1u << 17) ? std::min(8u, hw) : 1u;
    if (parts <= 1) {
        for (uint64_t i = 0; i < n; i++) RC(ok(i));
        return RT_OK;
    }
    std::vector<uint64_t> bad(parts, n);
    const uint64_t step = (n + parts - 1) / parts;
    auto scan = [&](unsigned c) {
        for (uint64_t i = (uint64_t)c * step, e = std::min(n, i + step); i < e; i++)
            if (ok(i) != RT_OK) {
                bad[c] = i;
                return;
            }
    };
    std::vector<std::thread> th;
    for (unsigned c = 1; c < parts; c++) {
        try {
            th.emplace_back(scan, c);
        } catch (const std::system_error &) {
            scan(c);
        }
    }
    scan(0);
    for (auto &t : th) t.join();
    const uint64_t first = *std::min_element(bad.begin(), bad.end());
    return first < n ? ok(first) : RT_OK;
}

static int validate_desc(const rt_scene_desc *d) {
    if (!d) return fail(RT_ERR_INVALID_ARGUMENT, "scene desc is null");
    if (d->abi_version != RT_B200_ABI_VERSION)
        return fail(RT_ERR_INVALID_ARGUMENT, "abi_version %u != %d", d->abi_version, RT_B200_ABI_VERSION);
    if ((d->n_spheres && !d->spheres) || (d->n_materials && !d->materials) || (d->n_textures && !d->textures) ||
        (d->n_images && !d->images) || (d->n_perlins && !d->perlins))
        return fail(RT_ERR_INVALID_ARGUMENT, "null array with non-zero count");
    if (d->n_quads && !d->quads) return fail(RT_ERR_INVALID_ARGUMENT, "null array with non-zero count");
    if (d->n_spheres >= (1ull << 27) || d->n_quads >= (1ull << 27))
        return fail(RT_ERR_INVALID_ARGUMENT, "too many hittables (%llu spheres, %llu quads)", (unsigned long long)d->n_spheres,
                    (unsigned long long)d->n_quads);
    if ((d->sphere_ids != nullptr) != (d->quad_ids != nullptr) && d->n_spheres && d->n_quads)
        return fail(RT_ERR_INVALID_ARGUMENT, "sphere_ids and quad_ids must be given together");
    auto quad_ok = [&](uint64_t i) -> int {
        const rt_quad &q = d->quads[i];
        if (q.material >= d->n_materials)
            return fail(RT_ERR_INVALID_ARGUMENT, "quad %llu: material %u out of range", (unsigned long long)i, q.material);
        for (int k = 0; k < 3; k++)
            if (!std::isfinite(q.q[k]) || !std::isfinite(q.u[k]) || !std::isfinite(q.v[k]))
                return fail(RT_ERR_INVALID_ARGUMENT, "quad %llu: non-finite geometry", (unsigned long long)i);
        return RT_OK;
    };
    auto sphere_ok = [&](uint64_t i) -> int {
        const rt_sphere &s = d->spheres[i];
        if (s.material >= d->n_materials)
            return fail(RT_ERR_INVALID_ARGUMENT, "sphere %llu: material %u out of range", (unsigned long long)i, s.material);
        if (!std::isfinite(s.cx) || !std::isfinite(s.cy) || !std::isfinite(s.cz) || !std::isfinite(s.r))
            return fail(RT_ERR_INVALID_ARGUMENT, "sphere %llu: non-finite geometry", (unsigned long long)i);
        return RT_OK;
    };
    auto material_ok = [&](uint64_t i) -> int {
        const rt_material &m = d->materials[i];
        if (m.kind > RT_MAT_DIFFUSE_LIGHT) return fail(RT_ERR_UNSUPPORTED, "material %u: unknown kind %u", (uint32_t)i, m.kind);
        if ((m.kind == RT_MAT_LAMBERTIAN || m.kind == RT_MAT_DIFFUSE_LIGHT) && m.texture >= d->n_textures)
            return fail(RT_ERR_INVALID_ARGUMENT, "material %u: texture %u out of range", (uint32_t)i, m.texture);
        return RT_OK;
    };
    auto texture_ok = [&](uint64_t i) -> int {
        const rt_texture &t = d->textures[i];
        if (t.kind > RT_TEX_NOISE) return fail(RT_ERR_UNSUPPORTED, "texture %u: unknown kind %u", (uint32_t)i, t.kind);
        if (t.kind == RT_TEX_NOISE && t.image >= d->n_perlins)
            return fail(RT_ERR_INVALID_ARGUMENT, "texture %u: perlin table %u out of range", (uint32_t)i, t.image);
        if (t.kind == RT_TEX_IMAGE && t.image >= d->n_images)
            return fail(RT_ERR_INVALID_ARGUMENT, "texture %u: image %u out of range", (uint32_t)i, t.image);
        return RT_OK;
    };
    RC(check_all(d->n_quads, quad_ok));
    RC(check_all(d->n_spheres, sphere_ok));
    RC(check_all(d->n_materials, material_ok));
    RC(check_all(d->n_textures, texture_ok));
    for (uint32_t i = 0; i < d->n_images; i++)
        if (d->images[i].w > 0 && d->images[i].h > 0 && !d->images[i].rgb16)
            return fail(RT_ERR_INVALID_ARGUMENT, "image %u: null texels", i);
    return RT_OK;
}

static void free_scene(rt_scene *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    // every render call returns after its stream has drained, so nothing is still using these
    cudaStream_t st = s->own_stream;
    scene_free(s->d_nodes, st), scene_free(s->d_sph, st), scene_free(s->d_mats, st), scene_free(s->d_meta, st);
    scene_free(s->d_images, st), scene_free(s->d_quads, st), scene_free(s->d_perlins, st), scene_free(s->d_chains, st);
    scene_free(s->d_raw, st), scene_free(s->d_raw_ids, st);
    for (auto p : s->d_texels) scene_free(p, st);
    scene_free(s->d_counter, st), scene_free(s->d_stats, st), scene_free(s->d_queue_count, st);
    for (auto e : s->events) cudaEventDestroy(e);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
}

static int upload_nodes(rt_scene *s) {
    if (!s->bvh.dev_nodes.empty())
        CU(cudaMemcpyAsync(s->d_nodes, s->bvh.dev_nodes.data(), s->bvh.dev_nodes.size() * sizeof(F4), cudaMemcpyHostToDevice, s->stream));
    return RT_OK;
}

// The device-independent half of a scene: copies of the hittables, the scene's centre / extent, the
// flattened BVH.  rt_render_multi builds it once and hands it to every device's scene_create_impl instead
// of letting N host threads build the same tree N times.
struct HostSceneParts {
    ScenePrims prims;
    double center[3] = {0, 0, 0}, extent90 = 0, surface_extent = 0;
    float origin_radius = 0;
    bool fixed_radius = false;
    FlatBvh bvh;
};
// Leaf-start chains (bvh_build.h) are built for scenes that can be staged in shared memory; RT_B200_LEAF_START=0
// turns them off (tuning knob / A-B measurement), RT_B200_LEAF_START_MAX moves the size limit.
static bool want_leaf_start(size_t n_prims) {
    return env_int("RT_B200_LEAF_START", 1) != 0 && n_prims <= (size_t)env_int("RT_B200_LEAF_START_MAX", 16384);
}

static int host_scene_parts(const rt_scene_desc *desc, HostSceneParts *h) {
    if (!load_scene_prims(desc, &h->prims))
        return fail(RT_ERR_INVALID_ARGUMENT, "sphere_ids / quad_ids are not a permutation of 0..n_hittables-1");
    compute_scene_center(h->prims, h->center, &h->extent90, &h->surface_extent);
    h->fixed_radius = desc->ray_origin_radius > 0;
    h->origin_radius = h->fixed_radius ? desc->ray_origin_radius : (float)(2.0 * h->extent90);
    build_flat_bvh(h->prims, h->origin_radius, env_int("RT_B200_MAX_LEAF", 4), &h->bvh, h->center, want_leaf_start(h->prims.size()));
    return RT_OK;
}

// Large sphere-only scenes get their tree from the GPU (bvh_device.cuh).  RT_B200_DEVICE_BVH: 0 never, 1 (default)
// from RT_B200_DEVICE_BVH_MIN spheres on (50 000), 2 whenever the scene has no quads (tests run the small parity
// scenes through the device builder that way).
static bool want_device_bvh(const rt_scene_desc *d) {
    const int mode = env_int("RT_B200_DEVICE_BVH", 1);
    if (mode == 0 || d->n_quads != 0 || d->n_spheres < 2) return false;
    return mode >= 2 || d->n_spheres >= (uint64_t)std::max(2, env_int("RT_B200_DEVICE_BVH_MIN", 50000));
}

// (Re)build the device tree of a device-built scene for s->origin_radius into the handle's buffers.
static int device_bvh_build(rt_scene *s) {
    cudaError_t e = device_build_bvh(s->d_raw, s->d_raw_ids, (uint32_t)s->n_prims, s->center, (double)s->origin_radius,
                                     // leaves of <= 2 spheres: measured on C4, 1 / 2 / 3 / 4 / 6 per leaf = 1329 / 1324 / 1307 / 1308 / 1289
                                     // Msamples/s (profiles/r02o); Morton-consecutive spheres make worse leaves than the SAH builder's
                                     std::max(1, std::min(env_int("RT_B200_MAX_LEAF", 2), RT_MAX_LEAF)), s->d_nodes, s->d_sph, s->d_meta,
                                     s->stream, &s->dev_stats);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY : RT_ERR_CUDA, "device BVH build: %s", cudaGetErrorString(e));
    return RT_OK;
}

// Uploads the spheres and builds the tree on the GPU.  Returns RT_OK with s->device_built = false when the scene turns
// out not to suit the device builder (too many huge spheres, a tree deeper than the traversal stack): the caller then
// uses the host builder.
static int scene_create_device_bvh(const rt_scene_desc *desc, int device, rt_scene *s) {
    const size_t n = desc->n_spheres;
    if (desc->sphere_ids) { // must be a permutation of 0..n-1 (load_scene_prims checks the same on the host path)
        std::vector<bool> seen(n, false);
        for (size_t i = 0; i < n; i++) {
            const uint32_t id = desc->sphere_ids[i];
            if (id >= n || seen[id]) return fail(RT_ERR_INVALID_ARGUMENT, "sphere_ids / quad_ids are not a permutation of 0..n_hittables-1");
            seen[id] = true;
        }
    }
    // the scene's centre and typical extent from a strided sample (exact medians of 1 M centres cost 13 ms and only
    // feed the padding rule, which needs a reference point, not THE median)
    {
        ScenePrims sample;
        const size_t stride = std::max<size_t>(1, n / 65536);
        for (size_t i = 0; i < n; i += stride) sample.spheres.push_back(desc->spheres[i]);
        sample.sphere_ids.resize(sample.spheres.size());
        compute_scene_center(sample, s->center, &s->extent90, nullptr);
    }
    s->fixed_radius = desc->ray_origin_radius > 0;
    s->origin_radius = s->fixed_radius ? desc->ray_origin_radius : (float)(2.0 * s->extent90);
    s->n_prims = n, s->has_quads = false;
    RC(scene_alloc(&s->d_raw, n * sizeof(rt_sphere), device, s->stream));
    CU(cudaMemcpyAsync(s->d_raw, desc->spheres, n * sizeof(rt_sphere), cudaMemcpyHostToDevice, s->stream));
    if (desc->sphere_ids) {
        RC(scene_alloc(&s->d_raw_ids, n * 4, device, s->stream));
        CU(cudaMemcpyAsync(s->d_raw_ids, desc->sphere_ids, n * 4, cudaMemcpyHostToDevice, s->stream));
    }
    RC(scene_alloc(&s->d_nodes, n * 64, device, s->stream)); // at most n pairs of nodes
    RC(scene_alloc(&s->d_sph, n * 16, device, s->stream));
    RC(scene_alloc(&s->d_meta, n * 8, device, s->stream));
    RC(device_bvh_build(s));
    if (s->dev_stats.n_large > BD_MAX_LARGE || s->dev_stats.max_depth + 2 > RT_LOCAL_STACK) { // not a scene for this builder
        scene_free(s->d_raw, s->stream), scene_free(s->d_raw_ids, s->stream), scene_free(s->d_nodes, s->stream);
        scene_free(s->d_sph, s->stream), scene_free(s->d_meta, s->stream);
        s->d_raw = nullptr, s->d_raw_ids = nullptr, s->d_nodes = nullptr, s->d_sph = nullptr, s->d_meta = nullptr;
        s->n_prims = 0;
        return RT_OK;
    }
    s->surface_extent = (double)__builtin_bit_cast(float, s->dev_stats.surface_extent);
    s->device_built = true;
    return RT_OK;
}

static int scene_create_impl(const rt_scene_desc *desc, int device, rt_scene *s, const HostSceneParts *pre = nullptr) {
    s->device = device;
    CU(cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device));
    int optin = 0;
    CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    s->smem_optin = (size_t)optin;
    CU(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    s->stream = s->own_stream;

    const bool timing = getenv("RT_B200_BVH_TIMING") != nullptr; // host-side phases of a scene creation
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[scene] %-14s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    if (!pre && want_device_bvh(desc)) {
        RC(scene_create_device_bvh(desc, device, s));
        lap(s->device_built ? "device bvh" : "device bvh (declined)");
    }
    if (s->device_built) {
        // nothing to do on the host
    } else if (pre) { // built once by the caller (rt_render_multi); every handle keeps its own copy (refits are per handle)
        s->prims = pre->prims, s->bvh = pre->bvh;
        s->center[0] = pre->center[0], s->center[1] = pre->center[1], s->center[2] = pre->center[2];
        s->extent90 = pre->extent90, s->surface_extent = pre->surface_extent;
        s->fixed_radius = pre->fixed_radius, s->origin_radius = pre->origin_radius;
        s->has_quads = !s->prims.quads.empty();
        lap("copy parts");
    } else {
        if (!load_scene_prims(desc, &s->prims))
            return fail(RT_ERR_INVALID_ARGUMENT, "sphere_ids / quad_ids are not a permutation of 0..n_hittables-1");
        lap("copy prims");
        s->has_quads = !s->prims.quads.empty();
        compute_scene_center(s->prims, s->center, &s->extent90, &s->surface_extent);
        s->fixed_radius = desc->ray_origin_radius > 0;
        s->origin_radius = s->fixed_radius ? desc->ray_origin_radius : (float)(2.0 * s->extent90);
        lap("scene centre");
        build_flat_bvh(s->prims, s->origin_radius, env_int("RT_B200_MAX_LEAF", 4), &s->bvh, s->center,
                       want_leaf_start(s->prims.size()));
        lap("bvh build");
    }
    if (!s->device_built) s->n_prims = s->prims.size();

    // materials: folded with their textures into 32-byte records — on the host, or (device-built scenes, which tend
    // to have one material per sphere) by a kernel from the raw arrays
    std::vector<F4> mats;
    const bool mats_on_device = s->device_built && desc->n_materials >= 4096;
    if (!mats_on_device) pack_materials(desc, &mats);
    lap("materials");

    // device nodes = the tree's nodes + the walk pairs of the leaf-start chains
    const size_t n_nodes = s->device_built ? 2 * (size_t)s->dev_stats.n_pairs : s->bvh.dev_nodes.size() / 2;
    const size_t n_slots = s->device_built ? s->n_prims : s->bvh.sph.size(), n_qslots = s->bvh.quad_prim.size();
    RC(scene_alloc(&s->d_quads, std::max<size_t>(1, n_qslots) * 16 * RT_QUAD_F4, device, s->stream));
    if (n_qslots)
        CU(cudaMemcpyAsync(s->d_quads, s->bvh.quad.data(), n_qslots * 16 * RT_QUAD_F4, cudaMemcpyHostToDevice, s->stream));
    if (!s->device_built) {
        RC(scene_alloc(&s->d_nodes, std::max<size_t>(1, n_nodes) * 32, device, s->stream));
        RC(scene_alloc(&s->d_sph, std::max<size_t>(1, n_slots) * 16, device, s->stream));
        RC(scene_alloc(&s->d_meta, std::max<size_t>(1, n_slots) * 8, device, s->stream));
        int rc = upload_nodes(s);
        if (rc != RT_OK) return rc;
        if (n_slots) {
            CU(cudaMemcpyAsync(s->d_sph, s->bvh.sph.data(), n_slots * 16, cudaMemcpyHostToDevice, s->stream));
            CU(cudaMemcpyAsync(s->d_meta, s->bvh.meta.data(), n_slots * 8, cudaMemcpyHostToDevice, s->stream));
        }
    }
    RC(scene_alloc(&s->d_mats, std::max<size_t>(1, (size_t)desc->n_materials) * 32, device, s->stream));
    if (mats_on_device) {
        rt_material *d_m = nullptr;
        rt_texture *d_t = nullptr;
        RC(scene_alloc(&d_m, (size_t)desc->n_materials * sizeof(rt_material), device, s->stream));
        RC(scene_alloc(&d_t, std::max<size_t>(1, desc->n_textures) * sizeof(rt_texture), device, s->stream));
        CU(cudaMemcpyAsync(d_m, desc->materials, (size_t)desc->n_materials * sizeof(rt_material), cudaMemcpyHostToDevice, s->stream));
        if (desc->n_textures)
            CU(cudaMemcpyAsync(d_t, desc->textures, (size_t)desc->n_textures * sizeof(rt_texture), cudaMemcpyHostToDevice, s->stream));
        bd_pack_materials<<<(desc->n_materials + 255) / 256, 256, 0, s->stream>>>(d_m, d_t, desc->n_materials, s->d_mats);
        CU(cudaGetLastError());
        scene_free(d_m, s->stream), scene_free(d_t, s->stream);
    } else if (!mats.empty()) {
        CU(cudaMemcpyAsync(s->d_mats, mats.data(), mats.size() * 16, cudaMemcpyHostToDevice, s->stream));
    }
    if (!s->bvh.chains.empty()) { // leaf start: chains, then the per-slot chain offsets of spheres and quads, in one buffer
        const size_t nc = s->bvh.chains.size(), total = nc + n_slots + n_qslots;
        RC(scene_alloc(&s->d_chains, total * 4, device, s->stream));
        CU(cudaMemcpyAsync(s->d_chains, s->bvh.chains.data(), nc * 4, cudaMemcpyHostToDevice, s->stream));
        if (n_slots) CU(cudaMemcpyAsync(s->d_chains + nc, s->bvh.sph_chain.data(), n_slots * 4, cudaMemcpyHostToDevice, s->stream));
        if (n_qslots)
            CU(cudaMemcpyAsync(s->d_chains + nc + n_slots, s->bvh.quad_chain.data(), n_qslots * 4, cudaMemcpyHostToDevice, s->stream));
        s->dev.chains = s->d_chains, s->dev.sph_chain = s->d_chains + nc, s->dev.quad_chain = s->d_chains + nc + n_slots;
        s->dev.n_chain_words = (uint32_t)nc;
    }

    // images: RGB16 -> (r, g, b, 0) uint16x4 so a texel is one 8-byte load
    std::vector<DevImage> imgs(desc->n_images);
    std::vector<std::vector<uint16_t>> staged(desc->n_images);
    for (uint32_t i = 0; i < desc->n_images; i++) {
        const rt_image &im = desc->images[i];
        imgs[i].w = im.w, imgs[i].h = im.h, imgs[i].texels = nullptr;
        if (im.w <= 0 || im.h <= 0) continue;
        const size_t n = (size_t)im.w * im.h;
        staged[i].resize(n * 4);
        for (size_t k = 0; k < n; k++) {
            staged[i][4 * k] = im.rgb16[3 * k], staged[i][4 * k + 1] = im.rgb16[3 * k + 1];
            staged[i][4 * k + 2] = im.rgb16[3 * k + 2], staged[i][4 * k + 3] = 0;
        }
        uint16_t *dp = nullptr;
        RC(scene_alloc(&dp, n * 8, device, s->stream));
        s->d_texels.push_back(dp);
        CU(cudaMemcpyAsync(dp, staged[i].data(), n * 8, cudaMemcpyHostToDevice, s->stream));
        imgs[i].texels = dp;
    }
    RC(scene_alloc(&s->d_images, std::max<size_t>(1, imgs.size()) * sizeof(DevImage), device, s->stream));
    if (!imgs.empty())
        CU(cudaMemcpyAsync(s->d_images, imgs.data(), imgs.size() * sizeof(DevImage), cudaMemcpyHostToDevice, s->stream));
    // Perlin tables (materials.go:195-200) as F4 gradients + the three permutations
    {
        std::vector<DevPerlin> pt(desc->n_perlins);
        for (uint32_t i = 0; i < desc->n_perlins; i++) {
            for (int k = 0; k < 256; k++) {
                pt[i].vec[k].x = desc->perlins[i].vec[k][0], pt[i].vec[k].y = desc->perlins[i].vec[k][1];
                pt[i].vec[k].z = desc->perlins[i].vec[k][2], pt[i].vec[k].w = 0;
            }
            memcpy(pt[i].perm_x, desc->perlins[i].perm_x, 256), memcpy(pt[i].perm_y, desc->perlins[i].perm_y, 256);
            memcpy(pt[i].perm_z, desc->perlins[i].perm_z, 256);
        }
        RC(scene_alloc(&s->d_perlins, std::max<size_t>(1, pt.size()) * sizeof(DevPerlin), device, s->stream));
        if (!pt.empty()) {
            CU(cudaMemcpyAsync(s->d_perlins, pt.data(), pt.size() * sizeof(DevPerlin), cudaMemcpyHostToDevice, s->stream));
            CU(cudaStreamSynchronize(s->stream)); // pt is a local
        }
    }
    RC(scene_alloc(&s->d_counter, sizeof(unsigned int), device, s->stream));
    RC(scene_alloc(&s->d_queue_count, RT_MAX_STAGES * sizeof(unsigned int), device, s->stream));
    RC(scene_alloc(&s->d_stats, RT_N_STATS * sizeof(unsigned long long), device, s->stream));
    CU(cudaStreamSynchronize(s->stream));

    s->dev.nodes = s->d_nodes, s->dev.sph = s->d_sph, s->dev.meta = s->d_meta, s->dev.mats = s->d_mats;
    s->dev.tex.images = s->d_images, s->dev.tex.perlins = s->d_perlins;
    s->dev.quads = s->d_quads, s->dev.n_quad_slots = (uint32_t)n_qslots;
    s->dev.root_ref = s->device_built ? s->dev_stats.root_ref : s->bvh.root_ref;
    s->dev.n_nodes = (uint32_t)n_nodes, s->dev.n_slots = (uint32_t)n_slots, s->dev.n_mats = desc->n_materials;
    s->n_images = desc->n_images, s->n_perlins = desc->n_perlins;
    s->dev.stack_depth = (s->device_built ? s->dev_stats.max_depth : s->bvh.max_depth) + 2;
    if (s->dev.stack_depth > RT_LOCAL_STACK) // cannot happen: the host builder balances the tree below depth 30, a deeper device-built tree was declined
        return fail(RT_ERR_INTERNAL, "BVH depth %u exceeds the traversal stack (%d)", s->dev.stack_depth - 2, RT_LOCAL_STACK);
    // 512-thread CTAs, two per SM: 32 warps at 64 registers.  Measured against 256 x 3 (24 warps at 80
    // registers): C2 +6 %, Cornell box +3 % (profiles/r01aj, r01ak); the kernels wait on fixed-latency
    // dependencies and shared-memory loads, so 8 more warps per SM buy more than 16 more registers.
    // 640 x 2 (40 warps, 48 registers, 200 B of spills) is 6.6 % slower again (profiles/r01al).
    // 1024 x 1 keeps the 32 warps and stages the scene ONCE per SM instead of twice: twice the shared-memory budget
    // (a scene + its leaf-start walk pairs that do not fit 2 x 113 KB fit 1 x 227 KB).  RT_B200_BLOCK = 0 (default):
    // 512 x 2 when the staging fits it, else 1024 x 1.
    const int want_block = env_int("RT_B200_BLOCK", 0);
    s->block = want_block == 256 || want_block == 512 || want_block == 1024 ? want_block : 512;
    s->minb = env_int("RT_B200_MINB", s->block == 256 ? 3 : s->block == 512 ? 2 : 1);
    if (s->block == 1024) s->minb = 1;
    const int want_pblock = env_int("RT_B200_PBLOCK", 0); // block size of the primary stage
    s->pblock = want_pblock == 256 || want_pblock == 512 || want_pblock == 1024 ? want_pblock : (s->block == 1024 ? 1024 : 512);
    {
        const char *kv = getenv("RT_B200_KERNEL");
        // default: two-stage ("split"); RT_B200_KERNEL=mega selects the one-stage megakernel
        s->use_split = kv ? std::string(kv) == "split" : true;
        s->n_stages = std::min(RT_MAX_STAGES, std::max(1, env_int("RT_B200_STAGES", 1)));
    }
    // the scene is staged in shared memory only if that still leaves room for every CTA the register
    // budget allows (minb per SM, 1 KB reserved per CTA); a larger scene is read through L1 instead,
    // which costs ~2 % on C2 — less than running fewer warps would
    int smem_sm = 0;
    CU(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    auto fits = [&](int block, int minb) {
        const size_t budget = std::min<size_t>(s->smem_optin, (size_t)smem_sm / (size_t)std::max(1, minb) - 1024);
        return smem_total_bytes(s->dev, block) <= budget;
    };
    const bool smem_ok = env_int("RT_B200_NO_SMEM", 0) == 0 && s->dev.stack_depth <= RT_LOCAL_STACK;
    s->use_smem = smem_ok && fits(s->block, s->minb) && fits(s->pblock, s->pblock == 1024 ? 1 : 2);
    if (smem_ok && !s->use_smem && want_block == 0 && want_pblock == 0 && fits(1024, 1)) {
        s->block = s->pblock = 1024, s->minb = 1;
        s->use_smem = true;
    }
    if (timing) CU(cudaStreamSynchronize(s->stream));
    lap("alloc + upload");
    return RT_OK;
}

static int rt_scene_create_impl(const rt_scene_desc *desc, int device, rt_scene **out, const HostSceneParts *pre = nullptr) {
    if (!out) return fail(RT_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    int rc = validate_desc(desc);
    if (rc != RT_OK) return rc;
    rc = select_device(device);
    if (rc != RT_OK) return rc;
    rt_scene *s = new rt_scene();
    try {
        rc = scene_create_impl(desc, device, s, pre);
    } catch (...) { // reported by guarded(); the half-built handle must not leak
        free_scene(s);
        throw;
    }
    if (rc != RT_OK) {
        free_scene(s);
        return rc;
    }
    *out = s;
    return RT_OK;
}

extern "C" int rt_scene_create(const rt_scene_desc *desc, int device, rt_scene **out) {
    return guarded([&] { return rt_scene_create_impl(desc, device, out); });
}


extern "C" void rt_scene_destroy(rt_scene *scene) { free_scene(scene); }

extern "C" int rt_scene_set_stream(rt_scene *scene, void *cuda_stream) {
    if (!scene) return fail(RT_ERR_INVALID_ARGUMENT, "scene is null");
    CU(cudaSetDevice(scene->device));
    CU(cudaStreamSynchronize(scene->stream));
    scene->stream = cuda_stream ? (cudaStream_t)cuda_stream : scene->own_stream;
    return RT_OK;
}

// Grow the BVH padding when ray origins lie outside the radius it was built for.
// `origin_dist` = largest distance of a ray origin from the scene centre; a ray from there can hit a
// primitive whose nearest point is origin_dist + surface_extent away.
static int ensure_origin_radius(rt_scene *s, double origin_dist) {
    const double needed = origin_dist + s->surface_extent;
    if (s->fixed_radius || needed <= s->origin_radius || s->n_prims == 0) return RT_OK;
    s->origin_radius = (float)(needed * 1.25);
    if (s->device_built) { // the same build with the larger padding: same Morton order, same topology, same buffers
        RC(device_bvh_build(s));
        if (2 * (size_t)s->dev_stats.n_pairs != s->dev.n_nodes || s->dev_stats.root_ref != s->dev.root_ref)
            return fail(RT_ERR_INTERNAL, "device BVH rebuild changed the topology");
        return RT_OK;
    }
    refit_flat_bvh(s->prims, s->origin_radius, &s->bvh);
    int rc = upload_nodes(s);
    if (rc != RT_OK) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return RT_OK;
}

static double dist_to_center(const rt_scene *s, const float *p) {
    double dx = p[0] - s->center[0], dy = p[1] - s->center[1], dz = p[2] - s->center[2];
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}

// ---------------------------------------------------------------------------------------------
// debug build (rt_debug.h): array sizes to the device before a call, violation counters back after it
// ---------------------------------------------------------------------------------------------
#if RT_DEBUG_CHECKS
static int dbg_begin(rt_scene *s, size_t queue_cap, size_t samples_cap, size_t n_pixels, size_t n_images, size_t n_perlins) {
    DbgBounds b;
    b.n_nodes = s->dev.n_nodes, b.n_slots = s->dev.n_slots, b.n_quad_slots = s->dev.n_quad_slots, b.n_mats = s->dev.n_mats;
    b.n_chain_words = s->dev.n_chain_words, b.stack_entries = s->use_smem ? s->dev.stack_depth : RT_LOCAL_STACK;
    b.n_images = (uint32_t)n_images, b.n_perlins = (uint32_t)n_perlins;
    b.queue_cap = queue_cap, b.samples_cap = samples_cap, b.n_pixels = n_pixels;
    if (env_int("RT_B200_DEBUG_TRIP", 0)) b.stack_entries = 1; // self-test: proves the checks are live (tests/test_gpu_debug_build.py)
    static const unsigned int zero[RT_DBG_N] = {0};
    CU(cudaMemcpyToSymbolAsync(g_dbg_bounds, &b, sizeof b, 0, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyToSymbolAsync(g_dbg_violations, zero, sizeof zero, 0, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream)); // `b` is a local
    return RT_OK;
}
static int dbg_end(rt_scene *s) {
    unsigned int v[RT_DBG_N];
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemcpyFromSymbol(v, g_dbg_violations, sizeof v, 0, cudaMemcpyDeviceToHost));
    for (int k = 0; k < RT_DBG_N; k++)
        if (v[k]) return fail(RT_ERR_INTERNAL, "debug build: bounds check '%s' failed %u time(s)", rt_dbg_names[k], v[k]);
    return RT_OK;
}
#else
static inline int dbg_begin(rt_scene *, size_t, size_t, size_t, size_t, size_t) { return RT_OK; }
static inline int dbg_end(rt_scene *) { return RT_OK; }
#endif

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (function, device) pair, not to a scene
// handle: with two live scenes of different size, setting it per handle would let the smaller scene
// lower the limit under the larger one (cudaErrorInvalidValue at its next launch).  It is therefore
// kept as a per-device, per-kernel high-water mark that is only ever raised.
template <class K>
static int ensure_dyn_smem(K kern, int device, size_t smem) {
    if (smem <= 48 * 1024) return RT_OK;
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> high;
    std::lock_guard<std::mutex> lock(mu);
    size_t &h = high[std::make_pair((const void *)kern, device)];
    if (smem > h) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h = smem;
    }
    return RT_OK;
}

template <int BLOCK, bool SMEM, bool COUNT, bool QUADS, bool FIRST>
static int launch_primary_t(rt_scene *s, const RenderParams &p) {
    auto kern = primary_stage_kernel<BLOCK, SMEM, COUNT, QUADS, FIRST>;
    if (s->primary_grid[COUNT][FIRST] == 0) {
        const size_t smem = SMEM ? smem_total_bytes(p.sc, BLOCK) : 0;
        RC(ensure_dyn_smem(kern, s->device, smem));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem));
        if (per_sm < 1) return fail(RT_ERR_CUDA, "primary stage does not fit on an SM (block %d, smem %zu)", BLOCK, smem);
        s->primary_grid[COUNT][FIRST] = s->sm_count * per_sm;
        s->primary_smem[COUNT][FIRST] = smem;
    }
    kern<<<s->primary_grid[COUNT][FIRST], BLOCK, s->primary_smem[COUNT][FIRST], s->stream>>>(p);
    CU(cudaGetLastError());
    return RT_OK;
}

template <int BLOCK, int MINB, bool SMEM, bool COUNT, bool QUADS, bool SPLIT = false>
static int launch_render_t(rt_scene *s, const RenderParams &p) {
    auto kern = render_kernel<BLOCK, MINB, SMEM, COUNT, QUADS, SPLIT>;
    if (s->grid_cache[COUNT] == 0) { // once per handle: opt in to the dynamic shared memory, size the grid
        const size_t smem = SMEM ? smem_total_bytes(p.sc, BLOCK) : 0;
        RC(ensure_dyn_smem(kern, s->device, smem));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem));
        if (per_sm < 1) return fail(RT_ERR_CUDA, "render kernel does not fit on an SM (block %d, smem %zu)", BLOCK, smem);
        s->grid_cache[COUNT] = s->sm_count * per_sm; // persistent: one wave of resident CTAs
        s->smem_cache[COUNT] = smem;
    }
    kern<<<s->grid_cache[COUNT], BLOCK, s->smem_cache[COUNT], s->stream>>>(p);
    CU(cudaGetLastError());
    return RT_OK;
}

// (block, min blocks/SM) instances: 512x2 = 32 warps at <= 64 registers (default), 1024x1 the same with one
// staging of the scene per SM, 256x3 = 24 warps at <= 85 registers.  RT_B200_BLOCK / RT_B200_MINB pick one (tuning knob).
template <bool SMEM, bool COUNT, bool QUADS>
static int launch_render_b(rt_scene *s, const RenderParams &p) {
    const int key = s->block * 10 + s->minb;
    switch (key) {
    case 10241: return launch_render_t<1024, 1, SMEM, COUNT, QUADS>(s, p);
    case 5122: return launch_render_t<512, 2, SMEM, COUNT, QUADS>(s, p);
    default: return launch_render_t<256, 3, SMEM, COUNT, QUADS>(s, p);
    }
}

// staged mode: n_stages coherent stage kernels (segment 0 from the camera, segment k from the queue of
// stage k-1, ping-pong buffers), then the megakernel on the survivors of the last stage
template <bool SMEM, bool COUNT, bool QUADS>
static int launch_split(rt_scene *s, const RenderParams &p0) {
    RenderParams p = p0;
    CU(cudaMemsetAsync(s->d_queue_count, 0, RT_MAX_STAGES * sizeof(unsigned int), s->stream));
    // queue layout: [buffer][o | d | t][queue_stride entries]; the two buffers are used in turn
    float4 *buf[2] = {p.queue_o, p.queue_o + 3 * p.queue_stride};
    for (int k = 0; k < s->n_stages; k++) {
        float4 *out = buf[k & 1];
        const float4 *in = buf[(k & 1) ^ 1];
        p.queue_o = out, p.queue_d = out + p.queue_stride, p.queue_t = out + 2 * p.queue_stride;
        p.queue_count = s->d_queue_count + k;
        p.in_o = in, p.in_d = in + p.queue_stride, p.in_t = in + 2 * p.queue_stride;
        p.in_count = k ? s->d_queue_count + k - 1 : nullptr;
        p.stage_depth = k;
        int rc;
        if (s->pblock == 1024)
            rc = k == 0 ? launch_primary_t<1024, SMEM, COUNT, QUADS, true>(s, p)
                        : launch_primary_t<1024, SMEM, COUNT, QUADS, false>(s, p);
        else if (s->pblock == 512)
            rc = k == 0 ? launch_primary_t<512, SMEM, COUNT, QUADS, true>(s, p)
                        : launch_primary_t<512, SMEM, COUNT, QUADS, false>(s, p);
        else
            rc = k == 0 ? launch_primary_t<256, SMEM, COUNT, QUADS, true>(s, p)
                        : launch_primary_t<256, SMEM, COUNT, QUADS, false>(s, p);
        if (rc != RT_OK) return rc;
    }
    p.stage_depth = s->n_stages; // the megakernel resumes the survivors of the last stage
    if (s->block == 1024) return launch_render_t<1024, 1, SMEM, COUNT, QUADS, true>(s, p);
    if (s->block == 512) return launch_render_t<512, 2, SMEM, COUNT, QUADS, true>(s, p);
    return launch_render_t<256, 3, SMEM, COUNT, QUADS, true>(s, p);
}

template <bool SMEM, bool COUNT>
static int launch_render_q(rt_scene *s, const RenderParams &p) {
    if (s->use_split) return s->has_quads ? launch_split<SMEM, COUNT, true>(s, p) : launch_split<SMEM, COUNT, false>(s, p);
    if (s->has_quads) return launch_render_b<SMEM, COUNT, true>(s, p);
    return launch_render_b<SMEM, COUNT, false>(s, p);
}

static int launch_render(rt_scene *s, const RenderParams &p, bool count) {
    if (s->use_smem) return count ? launch_render_q<true, true>(s, p) : launch_render_q<true, false>(s, p);
    return count ? launch_render_q<false, true>(s, p) : launch_render_q<false, false>(s, p);
}

static int check_camera(const rt_camera *c) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "camera is null");
    if (c->width < 1 || c->height < 1) return fail(RT_ERR_INVALID_ARGUMENT, "bad image size %dx%d", c->width, c->height);
    if ((int64_t)c->width * c->height > (1ll << 31) - 1) return fail(RT_ERR_INVALID_ARGUMENT, "image too large");
    if (c->max_depth < 0) return fail(RT_ERR_INVALID_ARGUMENT, "negative max_depth");
    return RT_OK;
}

static int scene_events(rt_scene *s, size_t n) {
    while (s->events.size() < n) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        s->events.push_back(e);
    }
    return RT_OK;
}

// The scanlines a call renders (rt_render_opts.row_*): image row = begin + k * step, k in [0, count).
struct RowSet {
    int begin, count, step;
};
static int row_set(const rt_camera *cam, const rt_render_opts *opts, RowSet *rs) {
    if (opts->row_count == 0) {
        rs->begin = 0, rs->count = cam->height, rs->step = 1;
        return RT_OK;
    }
    rs->begin = opts->row_begin, rs->count = opts->row_count, rs->step = opts->row_step;
    if (rs->count < 0 || rs->begin < 0 || rs->step < 1 ||
        (int64_t)rs->begin + (int64_t)(rs->count - 1) * rs->step >= (int64_t)cam->height)
        return fail(RT_ERR_INVALID_ARGUMENT, "row set (begin %d, count %d, step %d) outside the %d image rows", rs->begin,
                    rs->count, rs->step, cam->height);
    return RT_OK;
}

// Philox round keys of a seed -> the device's constant array (rt_rng.h).  Every launch that draws
// random numbers is preceded by this on its own stream; callers hold the device's workspace lock, so
// at most one seed is in use per device at a time.
static int upload_round_keys(uint64_t seed, cudaStream_t stream) {
    uint32_t rk[2 * RT_PHILOX_ROUNDS];
    philox_round_keys(seed, rk);
    CU(cudaMemcpyToSymbolAsync(c_philox_rk, rk, sizeof rk, 0, cudaMemcpyHostToDevice, stream));
    return RT_OK;
}

#define RT_PASS_PATHS (64u << 20) /* paths per pass: 1 GiB of float4 radiances + 3 GiB of survivor queue; measured
                                     on C2: 16 / 32 / 64 Mi paths per pass = 3377 / 3482 / 3557 Msamples/s (fewer tails) */

// Accumulate samples [sample_offset, +sample_count) of every pixel of the call's row set into
// d_accum (device, rows*W*3, compact).
static int render_accum(rt_scene *s, const rt_camera *cam, const rt_render_opts *opts, float *d_accum,
                        rt_stats *stats, uint32_t *launches) {
    RowSet rows;
    int rc = row_set(cam, opts, &rows);
    if (rc != RT_OK) return rc;
    const uint32_t n_pix = (uint32_t)cam->width * (uint32_t)rows.count;
    const int spp = opts->sample_count > 0 ? opts->sample_count : cam->spp;
    if (spp < 1) return fail(RT_ERR_INVALID_ARGUMENT, "sample count %d < 1", spp);
    if (opts->sample_offset < 0) return fail(RT_ERR_INVALID_ARGUMENT, "negative sample_offset");
    const bool count = (opts->flags & RT_FLAG_COUNT_WORK) != 0;
    // the camera may sit outside the radius the BVH was padded for
    const float lens = std::sqrt(cam->defocus_u[0] * cam->defocus_u[0] + cam->defocus_u[1] * cam->defocus_u[1] +
                                 cam->defocus_u[2] * cam->defocus_u[2]);
    rc = ensure_origin_radius(s, dist_to_center(s, cam->center) + 2.0 * lens);
    if (rc != RT_OK) return rc;

    const uint32_t budget = (uint32_t)std::max(1, env_int("RT_B200_PASS_PATHS", (int)RT_PASS_PATHS));
    const uint32_t pix_tile = std::min(n_pix, budget);
    const uint32_t spp_pass_max = std::max(1u, budget / pix_tile);
    const size_t need = (size_t)pix_tile * std::min<uint32_t>(spp_pass_max, (uint32_t)spp);
    Workspace &ws = g_ws[s->device];
    rc = ws_reserve(ws.samples, ws.samples_cap, need);
    if (rc == RT_OK && s->use_split) rc = ws_reserve(ws.queue, ws.queue_cap, (s->n_stages > 1 ? 6 : 3) * need);
    // per-pixel candidate lists for the camera rays (two-stage mode).  The walk costs 0.18 ms per C2 frame and saves
    // 0.012 ms per sample per pixel in the primary stage: worth it from ~15 samples per pixel on (at C1's 10 spp the
    // lists cost 2 %, profiles/r02l).  RT_B200_PIXEL_LISTS=0 turns them off (A/B)
    const bool use_lists = s->use_split && s->n_prims > 0 && spp >= env_int("RT_B200_PIXEL_LISTS_MIN_SPP", 16) &&
                           env_int("RT_B200_PIXEL_LISTS", 1) != 0;
    if (rc == RT_OK && use_lists) rc = ws_reserve(ws.lists, ws.lists_cap, (size_t)pix_tile * RT_LIST_WORDS);
    if (rc != RT_OK) return rc;
    CU(cudaMemsetAsync(s->d_stats, 0, RT_N_STATS * sizeof(unsigned long long), s->stream));
    RC(dbg_begin(s, need, need, n_pix, s->n_images, s->n_perlins));

    if (cam->max_depth <= 0) { // ray.go:33-35: every sample is black
        CU(cudaMemsetAsync(d_accum, 0, (size_t)n_pix * 3 * sizeof(float), s->stream));
        if (stats) stats->samples = (uint64_t)n_pix * (uint64_t)spp;
        return RT_OK;
    }
    rc = upload_round_keys(opts->seed, s->stream);
    if (rc != RT_OK) return rc;
    size_t n_ev = 2; // events 0 and 1 bracket the whole call
    RenderParams p;
    p.sc = s->dev;
    p.cam = make_dev_camera(*cam);
    p.seed = opts->seed;
    p.row_begin = (uint32_t)rows.begin, p.row_step = (uint32_t)rows.step;
    p.div_width = fast_div_magic((uint32_t)cam->width), p.div_spp = 0; // (image_pixel needs div_width: also in the candidate walk)
    p.samples = ws.samples;
    p.counter = s->d_counter;
    p.stats = s->d_stats;
    p.regen_min = (uint32_t)std::min(32, std::max(1, env_int("RT_B200_REGEN_MIN", s->use_split ? 1 : 8)));
    p.chunk = (uint32_t)std::min(1 << 16, std::max(32, env_int("RT_B200_CHUNK", (int)RT_CHUNK)));
    p.queue_o = ws.queue, p.queue_stride = need; // launch_split fills the per-stage pointers
    p.queue_d = p.queue_t = nullptr, p.queue_count = s->d_queue_count;
    p.in_o = p.in_d = p.in_t = nullptr, p.in_count = nullptr, p.stage_depth = 0;
    p.lists = nullptr;
    const bool balance = env_int("RT_B200_PASS_BALANCE", 0) != 0;
    uint64_t n_passes_total = 0;
    for (uint32_t pb = 0; pb < n_pix; pb += pix_tile) {
        const uint32_t np = std::min(pix_tile, n_pix - pb);
        // RT_B200_PASS_BALANCE=1 makes the passes equal (500 spp at 82 per pass = 7 passes of 71-72 instead of
        // 6 x 82 + 8); measured 0.3 % slower than the greedy split (profiles/r01y), so it is off by default
        const uint32_t n_passes = ((uint32_t)spp + spp_pass_max - 1) / spp_pass_max;
        if (use_lists) { // one beam walk per pixel of the tile, shared by all its samples
            p.pixel_begin = pb;
            p.lists = nullptr;
            if (s->has_quads) pixel_candidates_kernel<true><<<(np + 127) / 128, 128, 0, s->stream>>>(p, np, ws.lists);
            else pixel_candidates_kernel<false><<<(np + 127) / 128, 128, 0, s->stream>>>(p, np, ws.lists);
            CU(cudaGetLastError());
            p.lists = ws.lists;
            *launches += 1;
        }
        for (uint32_t k0 = 0, pass = 0, sp = 0; k0 < (uint32_t)spp; k0 += sp, pass++) {
            sp = balance ? ((uint32_t)spp - k0 + (n_passes - pass) - 1) / (n_passes - pass)
                         : std::min(spp_pass_max, (uint32_t)spp - k0);
            p.pixel_begin = pb;
            p.sample_begin = (uint32_t)opts->sample_offset + k0;
            p.spp_pass = sp;
            p.div_spp = fast_div_magic(sp), p.div_width = fast_div_magic((uint32_t)cam->width);
            p.total_paths = np * sp;
            CU(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned int), s->stream));
            rc = scene_events(s, n_ev + 2);
            if (rc != RT_OK) return rc;
            CU(cudaEventRecord(s->events[n_ev], s->stream));
            rc = launch_render(s, p, count);
            if (rc != RT_OK) return rc;
            CU(cudaEventRecord(s->events[n_ev + 1], s->stream));
            n_ev += 2;
            reduce_kernel<<<(np + 127) / 128, 128, 0, s->stream>>>(ws.samples, d_accum, pb, np, sp, k0 == 0);
            CU(cudaGetLastError());
            *launches += s->use_split ? 2 + s->n_stages : 2;
            n_passes_total++;
        }
    }
    RC(dbg_end(s));
    if (stats) {
        unsigned long long h[RT_N_STATS];
        CU(cudaMemcpyAsync(h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        stats->samples = (uint64_t)n_pix * (uint64_t)spp;
        stats->rays = h[0], stats->hits = h[1], stats->box_tests = h[2], stats->sphere_tests = h[3];
        stats->survivors = h[4];
        // HBM bytes of the call's work buffers: a 16-byte radiance record written and read per path, a 48-byte queue
        // entry written and read per survivor, the accumulator read + written per pass, the candidate lists written once
        stats->work_bytes = 32ull * stats->samples + 96ull * h[4] + 24ull * (uint64_t)n_pix * n_passes_total +
                            (use_lists ? (uint64_t)RT_LIST_WORDS * 4ull * n_pix : 0ull);
        float total = 0;
        for (size_t e = 2; e < n_ev; e += 2) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, s->events[e], s->events[e + 1]));
            total += ms;
        }
        stats->ms_megakernel = total, stats->megakernel_launches = (uint32_t)((n_ev - 2) / 2);
    }
    return RT_OK;
}

static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int rt_render_accum_device_impl(rt_scene *scene, const rt_camera *camera, const rt_render_opts *opts,
                                      float *d_accum, rt_stats *stats) {
    if (!scene || !opts || !d_accum) return fail(RT_ERR_INVALID_ARGUMENT, "null argument");
    int rc = check_camera(camera);
    if (rc != RT_OK) return rc;
    const double t0 = now_ms();
    CU(cudaSetDevice(scene->device));
    std::lock_guard<std::mutex> lock(g_ws[scene->device].mu);
    if (stats) memset(stats, 0, sizeof *stats);
    rc = scene_events(scene, 2);
    if (rc != RT_OK) return rc;
    CU(cudaEventRecord(scene->events[0], scene->stream));
    uint32_t launches = 0;
    rc = render_accum(scene, camera, opts, d_accum, stats, &launches);
    if (rc != RT_OK) return rc;
    CU(cudaEventRecord(scene->events[1], scene->stream));
    CU(cudaStreamSynchronize(scene->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, scene->events[0], scene->events[1]));
    if (stats) stats->ms_render = ms, stats->ms_total = (float)(now_ms() - t0), stats->kernel_launches = launches;
    return RT_OK;
}

extern "C" int rt_render_accum_device(rt_scene *scene, const rt_camera *camera, const rt_render_opts *opts,
                                      float *d_accum, rt_stats *stats) {
    return guarded([&] { return rt_render_accum_device_impl(scene, camera, opts, d_accum, stats); });
}


static int rt_resolve_device_impl(const float *d_accum, int32_t width, int32_t height, int32_t total_spp,
                                 int32_t device, void *cuda_stream, uint8_t *rgb_out) {
    if (!d_accum || !rgb_out || width < 1 || height < 1 || total_spp < 1) return fail(RT_ERR_INVALID_ARGUMENT, "bad argument");
    int rc = select_device(device);
    if (rc != RT_OK) return rc;
    if (device >= RT_MAX_DEVICES) return fail(RT_ERR_INVALID_ARGUMENT, "device ordinal too large");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const uint32_t n_pix = (uint32_t)width * (uint32_t)height;
    Workspace &ws = g_ws[device];
    std::lock_guard<std::mutex> lock(ws.mu);
    rc = ws_reserve(ws.rgb, ws.rgb_cap, (size_t)n_pix * 3);
    if (rc == RT_OK) rc = ws_reserve(ws.h_rgb, ws.h_rgb_cap, (size_t)n_pix * 3, true);
    if (rc != RT_OK) return rc;
    resolve_kernel<<<(n_pix + 255) / 256, 256, 0, st>>>(d_accum, ws.rgb, n_pix, 1.0f / (float)total_spp);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ws.h_rgb, ws.rgb, (size_t)n_pix * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(rgb_out, ws.h_rgb, (size_t)n_pix * 3);
    return RT_OK;
}

extern "C" int rt_resolve_device(const float *d_accum, int32_t width, int32_t height, int32_t total_spp,
                                 int32_t device, void *cuda_stream, uint8_t *rgb_out) {
    return guarded([&] { return rt_resolve_device_impl(d_accum, width, height, total_spp, device, cuda_stream, rgb_out); });
}


static int rt_render_impl(rt_scene *scene, const rt_camera *camera, const rt_render_opts *opts, uint8_t *rgb_out,
                         float *accum_out, rt_stats *stats) {
    if (!scene || !opts || !rgb_out) return fail(RT_ERR_INVALID_ARGUMENT, "null argument");
    int rc = check_camera(camera);
    if (rc != RT_OK) return rc;
    const double t0 = now_ms();
    CU(cudaSetDevice(scene->device));
    Workspace &ws = g_ws[scene->device];
    std::lock_guard<std::mutex> lock(ws.mu);
    if (stats) memset(stats, 0, sizeof *stats);
    RowSet rows;
    rc = row_set(camera, opts, &rows);
    if (rc != RT_OK) return rc;
    const uint32_t n_pix = (uint32_t)camera->width * (uint32_t)rows.count;
    if (n_pix == 0) return RT_OK;
    const int spp = opts->sample_count > 0 ? opts->sample_count : camera->spp;
    rc = ws_reserve(ws.accum, ws.accum_cap, (size_t)n_pix * 3);
    if (rc == RT_OK) rc = ws_reserve(ws.rgb, ws.rgb_cap, (size_t)n_pix * 3);
    if (rc == RT_OK) rc = ws_reserve(ws.h_rgb, ws.h_rgb_cap, (size_t)n_pix * 3, true);
    if (rc == RT_OK && accum_out) rc = ws_reserve(ws.h_accum, ws.h_accum_cap, (size_t)n_pix * 3, true);
    if (rc == RT_OK) rc = scene_events(scene, 2);
    if (rc != RT_OK) return rc;
    CU(cudaEventRecord(scene->events[0], scene->stream));
    uint32_t launches = 0;
    rc = render_accum(scene, camera, opts, ws.accum, stats, &launches);
    if (rc != RT_OK) return rc;
    resolve_kernel<<<(n_pix + 255) / 256, 256, 0, scene->stream>>>(ws.accum, ws.rgb, n_pix, 1.0f / (float)spp);
    launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(scene->events[1], scene->stream));
    CU(cudaMemcpyAsync(ws.h_rgb, ws.rgb, (size_t)n_pix * 3, cudaMemcpyDeviceToHost, scene->stream));
    if (accum_out)
        CU(cudaMemcpyAsync(ws.h_accum, ws.accum, (size_t)n_pix * 3 * sizeof(float), cudaMemcpyDeviceToHost, scene->stream));
    CU(cudaStreamSynchronize(scene->stream));
    memcpy(rgb_out, ws.h_rgb, (size_t)n_pix * 3);
    if (accum_out) memcpy(accum_out, ws.h_accum, (size_t)n_pix * 3 * sizeof(float));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, scene->events[0], scene->events[1]));
    if (stats) stats->ms_render = ms, stats->ms_total = (float)(now_ms() - t0), stats->kernel_launches = launches;
    return RT_OK;
}

extern "C" int rt_render(rt_scene *scene, const rt_camera *camera, const rt_render_opts *opts, uint8_t *rgb_out,
                         float *accum_out, rt_stats *stats) {
    return guarded([&] { return rt_render_impl(scene, camera, opts, rgb_out, accum_out, stats); });
}


// ---------------------------------------------------------------------------------------------
// one call, several GPUs (what a single host process such as the Go program needs)
// ---------------------------------------------------------------------------------------------
// Inside the library, on one host thread per device:
//  * sample-split (default): device k renders its share of the samples of every pixel into its own
//    FP32 accumulator; the accumulators are peer-copied to devices[0] and added in device order
//    (deterministic), devices[0] resolves.  Every (pixel, sample) keeps the Philox stream it has in
//    a single-GPU render, so only the FP32 summation order differs.
//  * tile-split (RT_FLAG_TILE_SPLIT): device k renders scanlines k, k+n, k+2n, ... at all samples
//    and resolves them itself; the host interleaves the rows.  No exchange between devices, and the
//    image is bit-identical to the single-GPU render.
// Direct peer access between every pair of the call's devices (NVLink / NVSwitch on a B200 box): peer copies then move
// device to device without staging.  Done once per pair and process; a pair that cannot be peers still works — the
// runtime stages those copies.
static void enable_peer_access(const std::vector<int> &devs) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, bool> done;
    std::lock_guard<std::mutex> lock(mu);
    for (int a : devs)
        for (int b : devs) {
            if (a == b || done.count(std::make_pair(a, b))) continue;
            done[std::make_pair(a, b)] = true;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) {
                cudaGetLastError();
                continue;
            }
            if (cudaSetDevice(a) == cudaSuccess) cudaDeviceEnablePeerAccess(b, 0);
            cudaGetLastError(); // cudaErrorPeerAccessAlreadyEnabled is fine
        }
}

static int rt_render_multi_impl(const rt_scene_desc *desc, const rt_camera *camera, const rt_render_opts *opts,
                               const int32_t *devices, int32_t n_devices, uint8_t *rgb_out, float *accum_out,
                               rt_stats *stats) {
    if (!desc || !opts || !rgb_out || !devices || n_devices < 1) return fail(RT_ERR_INVALID_ARGUMENT, "null argument");
    int rc = check_camera(camera);
    if (rc != RT_OK) return rc;
    const double t0 = now_ms();
    const int spp = opts->sample_count > 0 ? opts->sample_count : camera->spp;
    if (spp < 1) return fail(RT_ERR_INVALID_ARGUMENT, "sample count %d < 1", spp);
    RowSet rows;
    rc = row_set(camera, opts, &rows);
    if (rc != RT_OK) return rc;
    std::vector<int> devs(devices, devices + n_devices);
    {
        std::vector<int> sorted = devs;
        std::sort(sorted.begin(), sorted.end());
        if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end())
            return fail(RT_ERR_INVALID_ARGUMENT, "duplicate device in the device list");
        for (int d : sorted)
            if (d < 0 || d >= RT_MAX_DEVICES) return fail(RT_ERR_INVALID_ARGUMENT, "device %d out of range", d);
    }
    const bool tiles = (opts->flags & RT_FLAG_TILE_SPLIT) != 0;
    const size_t row_bytes = (size_t)camera->width * 3;
    const size_t n_acc = (size_t)rows.count * row_bytes;
    const uint32_t n_pix = (uint32_t)camera->width * (uint32_t)rows.count;
    if (n_pix == 0) return RT_OK;

    // workspaces are locked in ascending device order (no lock-order inversion between calls)
    std::vector<int> order = devs;
    std::sort(order.begin(), order.end());
    std::vector<std::unique_lock<std::mutex>> locks;
    for (int d : order) locks.emplace_back(g_ws[d].mu);

    struct Job {
        rt_scene *scene = nullptr;
        rt_render_opts o;
        rt_stats st;
        bool idle = false; // more devices than samples / rows
        int rc = RT_OK;
        std::string err;
    };
    std::vector<Job> jobs(n_devices);
    const int base = spp / n_devices, rem = spp % n_devices;
    // the BVH is built once, here, and copied into every device's handle
    rc = validate_desc(desc);
    if (rc != RT_OK) return rc;
    HostSceneParts parts;
    const bool per_device_build = want_device_bvh(desc); // every device builds its own tree (a few ms of GPU time)
    if (!per_device_build) {
        rc = host_scene_parts(desc, &parts);
        if (rc != RT_OK) return rc;
    }
    const bool timing = getenv("RT_B200_MULTI_TIMING") != nullptr; // host-side phases of the call, to stderr
    double t_lap = t0;
    auto lap = [&](const char *what) {
        if (!timing) return;
        const double t1 = now_ms();
        fprintf(stderr, "[multi] %-24s %7.2f ms\n", what, t1 - t_lap);
        t_lap = t1;
    };
    lap("validate + BVH build");
    // the root's buffers every device writes into or reads from, allocated before the device threads start
    const int root = devs[0];
    Workspace &w0 = g_ws[root];
    enable_peer_access(devs); // (leaves some other device current)
    RC(select_device(root));  // the buffers below belong to the root
    RC(ws_reserve(w0.rgb, w0.rgb_cap, (size_t)n_pix * 3));
    RC(ws_reserve(w0.h_rgb, w0.h_rgb_cap, (size_t)n_pix * 3, true));
    if (accum_out) RC(ws_reserve(w0.h_accum, w0.h_accum_cap, n_acc, true));
    unsigned char *root_rgb = nullptr;
    float *root_acc = nullptr;
    if (tiles) { // [sums (optional) | image]
        RC(ws_reserve(w0.gather, w0.gather_cap, (accum_out ? n_acc * sizeof(float) : 0) + (size_t)n_pix * 3));
        root_acc = reinterpret_cast<float *>(w0.gather);
        root_rgb = w0.gather + (accum_out ? n_acc * sizeof(float) : 0);
    } else {
        for (int k = 0; k < n_devices; k++) { // receive buffers of the reduction tree
            if (select_device(devs[k]) != RT_OK) continue; // reported by the device's thread
            RC(ws_reserve(g_ws[devs[k]].gather, g_ws[devs[k]].gather_cap, n_acc * sizeof(float)));
        }
    }
    lap("peer access + buffers");
    // a device thread that cannot be started joins the ones that were before the error is reported
    struct Joiner {
        std::vector<std::thread> v;
        ~Joiner() {
            for (auto &t : v)
                if (t.joinable()) t.join();
        }
    } joiner;
    std::vector<std::thread> &threads = joiner.v;
    for (int k = 0; k < n_devices; k++) {
        Job &j = jobs[k];
        j.o = *opts;
        j.o.device = devs[k];
        j.o.flags &= ~RT_FLAG_TILE_SPLIT;
        j.o.row_begin = rows.begin, j.o.row_count = rows.count, j.o.row_step = rows.step;
        if (tiles) { // rows k, k+n, ... of the caller's row set
            j.o.row_begin = rows.begin + k * rows.step;
            j.o.row_step = rows.step * n_devices;
            j.o.row_count = (rows.count - k + n_devices - 1) / n_devices;
            j.o.sample_count = spp;
            j.idle = j.o.row_count <= 0;
        } else {
            j.o.sample_count = base + (k < rem ? 1 : 0);
            j.o.sample_offset = opts->sample_offset + k * base + std::min(k, rem);
            j.idle = j.o.sample_count == 0;
        }
        memset(&j.st, 0, sizeof j.st);
        threads.emplace_back([&, k]() {
            Job &jj = jobs[k];
            if (jj.idle) return;
            jj.rc = guarded([&] { return rt_scene_create_impl(desc, devs[k], &jj.scene, per_device_build ? nullptr : &parts); });
            if (jj.rc == RT_OK) {
                Workspace &ws = g_ws[devs[k]];
                const size_t my_acc = (size_t)jj.o.row_count * row_bytes;
                jj.rc = ws_reserve(ws.accum, ws.accum_cap, my_acc);
                if (jj.rc == RT_OK && tiles) jj.rc = ws_reserve(ws.rgb, ws.rgb_cap, my_acc);
                uint32_t launches = 0;
                cudaStream_t st = jj.scene->stream;
                if (jj.rc == RT_OK) jj.rc = scene_events(jj.scene, 2);
                if (jj.rc == RT_OK) jj.rc = cudaEventRecord(jj.scene->events[0], st) == cudaSuccess ? RT_OK : RT_ERR_CUDA;
                if (jj.rc == RT_OK) jj.rc = render_accum(jj.scene, camera, &jj.o, ws.accum, &jj.st, &launches);
                if (jj.rc == RT_OK) {
                    cudaError_t e = cudaSuccess;
                    if (tiles) {
                        // resolve the device's own rows, then put them where they belong in the root's image: local
                        // row r is row k + r*n of the caller's row set — one strided peer copy over NVLink
                        const uint32_t my_pix = (uint32_t)(my_acc / 3);
                        resolve_kernel<<<(my_pix + 255) / 256, 256, 0, st>>>(ws.accum, ws.rgb, my_pix, 1.0f / (float)spp);
                        launches++;
                        e = cudaGetLastError();
                        if (e == cudaSuccess)
                            e = cudaMemcpy2DAsync(root_rgb + (size_t)k * row_bytes, (size_t)n_devices * row_bytes, ws.rgb, row_bytes,
                                                  row_bytes, (size_t)jj.o.row_count, cudaMemcpyDefault, st);
                        if (e == cudaSuccess && accum_out)
                            e = cudaMemcpy2DAsync(root_acc + (size_t)k * row_bytes, (size_t)n_devices * row_bytes * sizeof(float), ws.accum,
                                                  row_bytes * sizeof(float), row_bytes * sizeof(float), (size_t)jj.o.row_count,
                                                  cudaMemcpyDefault, st);
                    }
                    if (e == cudaSuccess) e = cudaEventRecord(jj.scene->events[1], st);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                    if (e != cudaSuccess) jj.rc = fail(RT_ERR_CUDA, "render failed on device %d: %s", devs[k], cudaGetErrorString(e));
                    else cudaEventElapsedTime(&jj.st.ms_render, jj.scene->events[0], jj.scene->events[1]);
                    jj.st.kernel_launches = launches;
                }
            }
            if (jj.rc != RT_OK) jj.err = rt_last_error(); // thread-local: carry it to the caller's thread
        });
    }
    for (auto &t : threads) t.join();
    lap("render (all devices)");
    auto cleanup = [&]() {
        for (auto &j : jobs)
            if (j.scene) free_scene(j.scene);
    };
    for (auto &j : jobs)
        if (j.rc != RT_OK) {
            const int code = j.rc;
            const std::string msg = j.err;
            cleanup();
            return fail(code, "%s", msg.c_str());
        }

    cudaError_t e = cudaSetDevice(root);
    cudaStream_t st0 = jobs[0].scene->stream;
    if (tiles) {
        // every device has put its rows into the root's buffers (peer copies above): one read-back
        if (e == cudaSuccess) e = cudaMemcpyAsync(w0.h_rgb, root_rgb, (size_t)n_pix * 3, cudaMemcpyDeviceToHost, st0);
        if (e == cudaSuccess && accum_out) e = cudaMemcpyAsync(w0.h_accum, root_acc, n_acc * sizeof(float), cudaMemcpyDeviceToHost, st0);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st0);
        if (e == cudaSuccess) {
            memcpy(rgb_out, w0.h_rgb, (size_t)n_pix * 3);
            if (accum_out) memcpy(accum_out, w0.h_accum, n_acc * sizeof(float));
        }
    } else {
        // Sum of the devices' accumulators on devices[0] by a binary tree: in round s device k (k a multiple of 2s)
        // receives the partial sum of device k+s by a peer copy over NVLink into its gather buffer and adds it —
        // log2(n) rounds, the copies of one round in flight together, instead of n-1 copies and adds queued one
        // after the other on the root.  The tree is fixed by n, so the result is deterministic.
        std::vector<int> active;
        for (int k = 0; k < n_devices; k++)
            if (!jobs[k].idle) active.push_back(k);
        std::vector<cudaEvent_t> done(active.size(), nullptr);
        for (size_t a = 0; a < active.size() && e == cudaSuccess; a++) {
            e = cudaSetDevice(devs[active[a]]);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[a], cudaEventDisableTiming);
        }
        for (size_t step = 1; step < active.size() && e == cudaSuccess; step *= 2)
            for (size_t a = 0; a + step < active.size() && e == cudaSuccess; a += 2 * step) {
                const int kr = active[a], ks = active[a + step];
                Workspace &wr = g_ws[devs[kr]];
                cudaStream_t sr = jobs[kr].scene->stream, ss = jobs[ks].scene->stream;
                e = cudaSetDevice(devs[ks]);
                if (e == cudaSuccess) e = cudaEventRecord(done[a + step], ss); // the sender's partial sum is complete
                if (e == cudaSuccess) e = cudaSetDevice(devs[kr]);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(sr, done[a + step], 0);
                if (e == cudaSuccess)
                    e = cudaMemcpyPeerAsync(wr.gather, devs[kr], g_ws[devs[ks]].accum, devs[ks], n_acc * sizeof(float), sr);
                if (e == cudaSuccess) {
                    add_kernel<<<(unsigned)((n_acc + 255) / 256), 256, 0, sr>>>(wr.accum, reinterpret_cast<const float *>(wr.gather), n_acc);
                    e = cudaGetLastError();
                }
            }
        if (e == cudaSuccess) e = cudaSetDevice(root);
        if (e == cudaSuccess) {
            resolve_kernel<<<(n_pix + 255) / 256, 256, 0, st0>>>(w0.accum, w0.rgb, n_pix, 1.0f / (float)spp);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(w0.h_rgb, w0.rgb, (size_t)n_pix * 3, cudaMemcpyDeviceToHost, st0);
            if (e == cudaSuccess && accum_out) e = cudaMemcpyAsync(w0.h_accum, w0.accum, n_acc * sizeof(float), cudaMemcpyDeviceToHost, st0);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st0);
            if (e == cudaSuccess) {
                memcpy(rgb_out, w0.h_rgb, (size_t)n_pix * 3);
                if (accum_out) memcpy(accum_out, w0.h_accum, n_acc * sizeof(float));
            }
        }
        for (auto ev : done)
            if (ev) cudaEventDestroy(ev);
    }
    lap("exchange + read-back");
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (auto &j : jobs) {
            stats->samples += j.st.samples, stats->rays += j.st.rays, stats->hits += j.st.hits;
            stats->box_tests += j.st.box_tests, stats->sphere_tests += j.st.sphere_tests;
            stats->survivors += j.st.survivors, stats->work_bytes += j.st.work_bytes;
            stats->kernel_launches += j.st.kernel_launches, stats->megakernel_launches += j.st.megakernel_launches;
            stats->ms_render = std::max(stats->ms_render, j.st.ms_render);
            stats->ms_megakernel = std::max(stats->ms_megakernel, j.st.ms_megakernel);
        }
        stats->ms_total = (float)(now_ms() - t0);
    }
    cleanup();
    lap("destroy scenes");
    if (rc != RT_OK) return rc;
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, "multi-device exchange failed: %s", cudaGetErrorString(e));
    return RT_OK;
}

extern "C" int rt_render_multi(const rt_scene_desc *desc, const rt_camera *camera, const rt_render_opts *opts,
                               const int32_t *devices, int32_t n_devices, uint8_t *rgb_out, float *accum_out,
                               rt_stats *stats) {
    return guarded([&] { return rt_render_multi_impl(desc, camera, opts, devices, n_devices, rgb_out, accum_out, stats); });
}


template <int BLOCK, bool QUADS>
static int launch_trace_t(rt_scene *s, const float *d_o, const float *d_d, int64_t n, float tmin, float tmax,
                          int32_t *d_id, float *d_t) {
    const int grid = (int)std::min<int64_t>((n + BLOCK - 1) / BLOCK, (int64_t)s->sm_count * 16);
    const size_t smem = smem_total_bytes(s->dev, BLOCK);
    if (s->use_smem && smem <= s->smem_optin) {
        auto kern = trace_kernel<BLOCK, true, QUADS>;
        RC(ensure_dyn_smem(kern, s->device, smem));
        kern<<<std::min(grid, s->sm_count * 2), BLOCK, smem, s->stream>>>(s->dev, d_o, d_d, n, tmin, tmax, d_id, d_t);
    } else {
        trace_kernel<BLOCK, false, QUADS><<<grid, BLOCK, 0, s->stream>>>(s->dev, d_o, d_d, n, tmin, tmax, d_id, d_t);
    }
    CU(cudaGetLastError());
    return RT_OK;
}

static int rt_trace_impl(rt_scene *scene, const float *origins, const float *dirs, int64_t n, float tmin, float tmax,
                        int32_t *id_out, float *t_out) {
    if (!scene || n < 0 || (n && (!origins || !dirs || !id_out || !t_out))) return fail(RT_ERR_INVALID_ARGUMENT, "bad argument");
    if (n == 0) return RT_OK;
    CU(cudaSetDevice(scene->device));
    double far = 0;
    for (int64_t i = 0; i < n; i++) far = std::max(far, dist_to_center(scene, origins + 3 * i));
    if (!std::isfinite(far)) return fail(RT_ERR_INVALID_ARGUMENT, "non-finite ray origin");
    int rc = ensure_origin_radius(scene, far);
    if (rc != RT_OK) return rc;
    float *d_o = nullptr, *d_d = nullptr, *d_t = nullptr;
    int32_t *d_id = nullptr;
    auto cleanup = [&]() { cudaFree(d_o), cudaFree(d_d), cudaFree(d_t), cudaFree(d_id); };
    cudaError_t e = cudaMalloc(&d_o, (size_t)n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_d, (size_t)n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_t, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_id, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_o, origins, (size_t)n * 12, cudaMemcpyHostToDevice, scene->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_d, dirs, (size_t)n * 12, cudaMemcpyHostToDevice, scene->stream);
    if (e != cudaSuccess) {
        cleanup();
        return fail(e == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY : RT_ERR_CUDA, "rt_trace setup: %s", cudaGetErrorString(e));
    }
    rc = dbg_begin(scene, 0, 0, 0, scene->n_images, scene->n_perlins);
    if (rc == RT_OK)
        rc = scene->has_quads ? launch_trace_t<256, true>(scene, d_o, d_d, n, tmin, tmax, d_id, d_t)
                              : launch_trace_t<256, false>(scene, d_o, d_d, n, tmin, tmax, d_id, d_t);
    if (rc == RT_OK) rc = dbg_end(scene);
    if (rc == RT_OK) {
        e = cudaMemcpyAsync(id_out, d_id, (size_t)n * 4, cudaMemcpyDeviceToHost, scene->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(t_out, d_t, (size_t)n * 4, cudaMemcpyDeviceToHost, scene->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(scene->stream);
        if (e != cudaSuccess) rc = fail(RT_ERR_CUDA, "rt_trace: %s", cudaGetErrorString(e));
    }
    cleanup();
    return rc;
}

extern "C" int rt_trace(rt_scene *scene, const float *origins, const float *dirs, int64_t n, float tmin, float tmax,
                        int32_t *id_out, float *t_out) {
    return guarded([&] { return rt_trace_impl(scene, origins, dirs, n, tmin, tmax, id_out, t_out); });
}


static int rt_primary_rays_impl(const rt_camera *camera, const rt_render_opts *opts, int64_t pixel_begin, int64_t n_pixels,
                               float *origins_out, float *dirs_out) {
    if (!opts || !origins_out || !dirs_out) return fail(RT_ERR_INVALID_ARGUMENT, "null argument");
    int rc = check_camera(camera);
    if (rc != RT_OK) return rc;
    const int spp = opts->sample_count > 0 ? opts->sample_count : camera->spp;
    if (pixel_begin < 0 || n_pixels < 0 || pixel_begin + n_pixels > (int64_t)camera->width * camera->height || spp < 1)
        return fail(RT_ERR_INVALID_ARGUMENT, "pixel/sample range out of bounds");
    rc = select_device(opts->device);
    if (rc != RT_OK) return rc;
    const int64_t n = n_pixels * spp;
    if (n == 0) return RT_OK;
    if (opts->device >= RT_MAX_DEVICES) return fail(RT_ERR_INVALID_ARGUMENT, "device ordinal too large");
    std::lock_guard<std::mutex> lock(g_ws[opts->device].mu); // the round keys are per device
    rc = upload_round_keys(opts->seed, 0);
    if (rc != RT_OK) return rc;
    float *d_o = nullptr, *d_d = nullptr;
    cudaError_t e = cudaMalloc(&d_o, (size_t)n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_d, (size_t)n * 12);
    if (e == cudaSuccess) {
        primary_kernel<<<(unsigned)((n + 255) / 256), 256>>>(make_dev_camera(*camera), opts->seed, (uint32_t)pixel_begin,
                                                             (uint32_t)opts->sample_offset, (uint32_t)spp, n, d_o, d_d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(origins_out, d_o, (size_t)n * 12, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dirs_out, d_d, (size_t)n * 12, cudaMemcpyDeviceToHost);
    cudaFree(d_o), cudaFree(d_d);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, "rt_primary_rays: %s", cudaGetErrorString(e));
    return RT_OK;
}

extern "C" int rt_primary_rays(const rt_camera *camera, const rt_render_opts *opts, int64_t pixel_begin, int64_t n_pixels,
                               float *origins_out, float *dirs_out) {
    return guarded([&] { return rt_primary_rays_impl(camera, opts, pixel_begin, n_pixels, origins_out, dirs_out); });
}


// ---------------------------------------------------------------------------------------------
// host-only helpers
// ---------------------------------------------------------------------------------------------
// Camera.init, camera.go:128-166 (float32, unfused: this file is compiled with -ffp-contract=off).
extern "C" int rt_camera_from_options(const rt_camera_options *o, rt_camera *c) {
    if (!o || !c) return fail(RT_ERR_INVALID_ARGUMENT, "null argument");
    if (o->image_width < 1 || !(o->aspect_ratio > 0)) return fail(RT_ERR_INVALID_ARGUMENT, "bad image width / aspect");
    const V3 look_from = v3(o->look_from[0], o->look_from[1], o->look_from[2]);
    const V3 look_at = v3(o->look_at[0], o->look_at[1], o->look_at[2]);
    const V3 vup = v3(o->vup[0], o->vup[1], o->vup[2]);
    const float image_width = (float)o->image_width;               // camera.go:107
    const V3 center = look_from;                                   // :130
    const V3 dist = look_from - look_at;                           // :132
    const float h = (float)tan((double)(o->fov_radians / 2.0f));   // :134
    const float viewport_height = 2.0f * h * o->focus_dist;        // :135
    float image_height = (float)(floor((double)image_width) / (double)o->aspect_ratio); // :137
    if (image_height < 1) image_height = 1;
    const float viewport_width = viewport_height * (image_width / image_height); // :141
    const V3 w = unit(dist);
    auto cross = [](V3 l, V3 r) { return v3(l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x); };
    const V3 u = unit(cross(vup, w));
    const V3 v = cross(w, u);
    const V3 viewport_u = u * viewport_width;
    const V3 viewport_v = v * -viewport_height;
    const V3 du = viewport_u * (1 / image_width);
    const V3 dv = viewport_v * (1 / image_height);
    V3 ul = center;
    ul = ul - w * o->focus_dist;
    ul = ul - viewport_u * 0.5f;
    ul = ul - viewport_v * 0.5f;
    const V3 pixel00 = ul + (du + dv) * 0.5f;
    const float defocus_radius = o->focus_dist * (float)tan((double)(o->defocus_angle_radians / 2.0f));
    const V3 disk_u = u * defocus_radius, disk_v = v * defocus_radius;
    auto put = [](float *dst, V3 s) { dst[0] = s.x, dst[1] = s.y, dst[2] = s.z; };
    c->width = (int32_t)image_width, c->height = (int32_t)image_height;
    c->spp = o->spp, c->max_depth = o->max_depth;
    put(c->center, center), put(c->pixel00, pixel00), put(c->pixel_du, du), put(c->pixel_dv, dv);
    put(c->defocus_u, disk_u), put(c->defocus_v, disk_v);
    c->defocus_angle = o->defocus_angle_radians;
    c->background[0] = o->background[0], c->background[1] = o->background[1], c->background[2] = o->background[2];
    return RT_OK;
}

extern "C" int rt_scene_bvh_info(const rt_scene *s, rt_bvh_info *out) {
    if (!s || !out) return fail(RT_ERR_INVALID_ARGUMENT, "null argument");
    out->n_nodes = s->bvh.nodes.size() / 2, out->n_slots = s->bvh.sph.size();
    out->max_depth = s->bvh.max_depth, out->in_shared_memory = s->use_smem ? 1 : 0;
    out->root_ref = s->bvh.root_ref, out->built_on_device = 0;
    out->box_pad_min = s->bvh.pad_min, out->box_pad_max = s->bvh.pad_max;
    if (s->device_built) {
        out->n_nodes = 2 * (uint64_t)s->dev_stats.n_pairs, out->n_slots = s->n_prims;
        out->max_depth = s->dev_stats.max_depth, out->root_ref = s->dev_stats.root_ref;
        out->built_on_device = 1;
        out->box_pad_min = __builtin_bit_cast(float, s->dev_stats.pad_min), out->box_pad_max = __builtin_bit_cast(float, s->dev_stats.pad_max);
    }
    return RT_OK;
}

static int rt_scene_bvh_copy_impl(const rt_scene *s, uint32_t *nodes_out, int32_t *slot_ids_out) {
    if (!s) return fail(RT_ERR_INVALID_ARGUMENT, "null argument");
    if (s->device_built) { // read the tree back: boxes as the kernels test them, [c - h, c + h]
        CU(cudaSetDevice(s->device));
        const size_t nn = 2 * (size_t)s->dev_stats.n_pairs;
        if (nodes_out && nn) {
            std::vector<F4> dn(2 * nn);
            CU(cudaMemcpy(dn.data(), s->d_nodes, dn.size() * sizeof(F4), cudaMemcpyDeviceToHost));
            F4 *o = reinterpret_cast<F4 *>(nodes_out);
            for (size_t i = 0; i < nn; i++) {
                const F4 c = dn[2 * i], h = dn[2 * i + 1];
                o[2 * i] = F4{c.x - h.x, c.y - h.y, c.z - h.z, c.w}, o[2 * i + 1] = F4{c.x + h.x, c.y + h.y, c.z + h.z, 0.0f};
            }
        }
        if (slot_ids_out && s->n_prims) {
            std::vector<I2> meta(s->n_prims);
            CU(cudaMemcpy(meta.data(), s->d_meta, meta.size() * sizeof(I2), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < meta.size(); i++) slot_ids_out[i] = meta[i].x;
        }
        return RT_OK;
    }
    if (nodes_out && !s->bvh.nodes.empty()) memcpy(nodes_out, s->bvh.nodes.data(), s->bvh.nodes.size() * sizeof(F4));
    if (slot_ids_out)
        for (size_t i = 0; i < s->bvh.meta.size(); i++) slot_ids_out[i] = s->bvh.meta[i].x;
    return RT_OK;
}

extern "C" int rt_scene_bvh_copy(const rt_scene *s, uint32_t *nodes_out, int32_t *slot_ids_out) {
    return guarded([&] { return rt_scene_bvh_copy_impl(s, nodes_out, slot_ids_out); });
}

