/* rt_b200.h — C ABI of librt_b200.so, the B200 (sm_100a) path-tracing hot path that sits
 * behind raytracer-go's `Camera.Render(world, writer)`.
 *
 * Citations `file:line` are relative to the reference checkout (TwFlem/raytracer-go).
 * The reference has no FFI of its own (pure Go); each entry point below names the Go
 * function(s) it replaces.  INTEGRATION.md shows the cgo file that binds this header.
 *
 * Conventions
 *   - every function returns RT_OK (0) or a negative rt_status; rt_last_error() gives the
 *     thread-local message.  Nothing in the library calls abort()/exit().
 *   - the caller owns every buffer it passes; the library copies what it needs before the
 *     call returns and keeps no caller pointers (cgo rule: no Go pointers retained by C).
 *   - all arithmetic on the path is IEEE-754 binary32, unfused, in the reference's operation
 *     order (the gc compiler does not contract on amd64); f64 only where the reference
 *     round-trips through math.* (sqrt / pow / acos / atan2 / tan / floor).
 *   - a handle is not re-entrant (one render at a time per handle); distinct handles are
 *     independent.  Any OS thread may call in; each call selects its own device.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 5

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID_ARGUMENT = -1, /* null pointer, bad index, bad size                     */
    RT_ERR_CUDA = -2,             /* a CUDA runtime call failed (message has the cudaError) */
    RT_ERR_NO_DEVICE = -3,        /* no usable sm_100 device: the library has NO CPU path   */
    RT_ERR_OUT_OF_MEMORY = -4,
    RT_ERR_UNSUPPORTED = -5,      /* e.g. a material/texture kind this build does not know  */
    RT_ERR_INTERNAL = -6          /* a C++ exception reached the boundary (never thrown through) */
} rt_status;

/* materials.go:19-21,44-47,77-79,297-299 — the Material implementations. */
typedef enum rt_material_kind {
    RT_MAT_LAMBERTIAN = 0,   /* materials.go:33-42  */
    RT_MAT_METAL = 1,        /* materials.go:60-75  */
    RT_MAT_DIELECTRIC = 2,   /* materials.go:91-119 */
    RT_MAT_DIFFUSE_LIGHT = 3 /* materials.go:297-313 (emit only) */
} rt_material_kind;

/* materials.go:121-125,151-153,165-167 — the Texture implementations. */
typedef enum rt_texture_kind {
    RT_TEX_SOLID = 0,   /* materials.go:155-157 */
    RT_TEX_CHECKER = 1, /* materials.go:127-137 */
    RT_TEX_IMAGE = 2,   /* materials.go:175-193 */
    RT_TEX_NOISE = 3    /* materials.go:280-295: marble from Perlin turbulence */
} rt_texture_kind;

/* hittables.go:78-83 `Sphere{Center, Radius, Material}`; the bounding box (hittables.go:91) is
 * rebuilt by the library.  Index in the array == insertion order in World.hittables
 * (hittables.go:48-53) == the object ID rt_trace reports. */
typedef struct rt_sphere {
    float cx, cy, cz;
    float r;
    uint32_t material; /* index into rt_scene_desc.materials */
} rt_sphere;

/* hittables.go:138-147 `Quad{Q, u, v, material}`: the parallelogram Q + a*u + b*v, 0 <= a,b <= 1.
 * w, normal, D and the padded box (hittables.go:149-165) are derived by the library. */
typedef struct rt_quad {
    float q[3];
    float u[3];
    float v[3];
    uint32_t material;
} rt_quad;

typedef struct rt_material {
    uint32_t kind;    /* rt_material_kind                                              */
    float albedo[3];  /* Metal.albedo (materials.go:45)                                */
    float fuzz;       /* Metal.fuzz   (materials.go:46)                                */
    float ior;        /* Dielectric.refractiveIndex (materials.go:78)                  */
    uint32_t texture; /* Lambertian.albedo / DiffuseLight.emit: index into textures    */
} rt_material;

typedef struct rt_texture {
    uint32_t kind;  /* rt_texture_kind                                                     */
    float a[3];     /* SolidColor.albedo (materials.go:152) or Checkered.even (:123)       */
    float b[3];     /* Checkered.odd (:124)                                                */
    float scale;    /* Checkered.scale (:122) or NoiseTexture.scale (:282)                 */
    uint32_t image; /* ImageTexture: index into images; NoiseTexture: index into perlins   */
    float oob[3];   /* ImageTexture: colour Go's img.At() returns outside Bounds(), i.e.
                       img.At(W,0).RGBA()/65535 evaluated once by the bridge */
} rt_texture;

/* ImageTexture.img pre-decoded by the Go side: texel (i,j) = img.At(i,j).RGBA() r,g,b as
 * 16-bit values (materials.go:186-189), row-major, 3 x uint16 per texel. */
typedef struct rt_image {
    int32_t w, h;
    const uint16_t *rgb16;
} rt_image;

/* Perlin's tables (materials.go:195-200, built by NewPerlin :202-216): 256 gradient vectors in
 * [-1,1)^3 (not normalised) and three permutations of 0..255.  The bridge copies them from the Go
 * object, so the device evaluates the very noise field the Go scene holds. */
typedef struct rt_perlin {
    float vec[256][3];
    uint8_t perm_x[256], perm_y[256], perm_z[256];
} rt_perlin;

typedef struct rt_scene_desc {
    uint32_t abi_version; /* RT_B200_ABI_VERSION */
    uint32_t reserved;
    const rt_sphere *spheres;
    uint64_t n_spheres;
    const rt_material *materials;
    uint32_t n_materials;
    const rt_texture *textures;
    uint32_t n_textures;
    const rt_image *images;
    uint32_t n_images;
    /* Sizes the conservative padding of the device BVH boxes: box culling never drops a primitive the
     * reference's float32 Hit would accept for any ray that starts within this distance (world units)
     * of the surface it hits.  0 = derive it from the scene (twice the 90th-percentile extent) and enlarge
     * it as needed to cover the camera of an rt_render call or the origins of an rt_trace batch: camera rays
     * and traced batches are then always exact; a SECONDARY ray of rt_render is exact when it starts within
     * the radius of the surface it hits — from farther away (e.g. off a huge ground sphere, 100 units out) a
     * hit that exists only through the rounding noise of the reference's float32 discriminant (a geometric
     * miss by up to 20 u |o-c|^2 / 2r) can be culled.  > 0 = a fixed envelope (use it for scenes much larger
     * than the distances at which float32 sphere tests are still meaningful; measured
     * agreement beyond the envelope: profiles/r02j_c4_parity_by_distance.txt). */
    float ray_origin_radius;
    /* Quads (Quad/Box, hittables.go:138-216).  Object IDs: the position of a
     * hittable in World.hittables (hittables.go:48-53), which World.Hit breaks exact ties by.  When
     * sphere_ids / quad_ids are NULL, spheres are objects 0..n_spheres-1 and quads follow; otherwise
     * the two arrays together must be a permutation of 0..n_spheres+n_quads-1. */
    const rt_quad *quads;
    uint64_t n_quads;
    const uint32_t *sphere_ids;
    const uint32_t *quad_ids;
    const rt_perlin *perlins; /* NoiseTexture.perlin, indexed by rt_texture.image */
    uint32_t n_perlins;
    uint32_t reserved2;
} rt_scene_desc;

/* The derived Camera fields of camera.go:23-52 as computed by Camera.init (camera.go:128-166).
 * The Go bridge copies them from its own *Camera; rt_camera_from_options computes them for
 * callers that only have the option values. */
typedef struct rt_camera {
    int32_t width;      /* int(imageWidth)  camera.go:181 */
    int32_t height;     /* int(imageHeight) camera.go:182 */
    int32_t spp;        /* samplesPerPixel  camera.go:29  */
    int32_t max_depth;  /* bounceDepth      camera.go:30  */
    float center[3];    /* camera.go:130 */
    float pixel00[3];   /* camera.go:160-161 */
    float pixel_du[3];  /* camera.go:150-151 */
    float pixel_dv[3];  /* camera.go:152-153 */
    float defocus_u[3]; /* camera.go:164 */
    float defocus_v[3]; /* camera.go:165 */
    float defocus_angle; /* defocusAngleRadians: only its sign is read on the path, camera.go:279 */
    float background[3]; /* camera.go:51, ray.go:53 */
} rt_camera;

/* The CameraOpt values (camera.go:54-102) with NewCamera's defaults (camera.go:105-117). */
typedef struct rt_camera_options {
    float aspect_ratio;
    int32_t image_width;
    int32_t spp;
    int32_t max_depth;
    float fov_radians;           /* WithFOVDegrees stores ToRadians(deg), camera.go:70 */
    float defocus_angle_radians; /* camera.go:88 */
    float focus_dist;
    float look_from[3], look_at[3], vup[3];
    float background[3];
} rt_camera_options;

typedef struct rt_render_opts {
    uint64_t seed;         /* Philox4x32-7 key; the reference is clock-seeded (camera.go:170)     */
    int32_t device;        /* CUDA ordinal                                                        */
    int32_t sample_offset; /* first global sample index rendered by this call (sample-split)      */
    int32_t sample_count;  /* samples per pixel rendered by this call; 0 = camera.spp             */
    int32_t flags;         /* RT_FLAG_*                                                           */
    /* Row set (tile-split): this call renders only the scanlines
     * row_begin + k*row_step, k in [0, row_count).  row_count == 0 = the whole image (row_begin and
     * row_step are then ignored).  Every output buffer of the call (rgb_out, accum_out, d_accum) is
     * COMPACT: row_count rows of `width` pixels, in increasing k.  Pixels keep the Philox stream of
     * their position in the full image, so a row set is bit-identical to the same rows of the
     * full render.  row_step > 1 interleaves the devices' rows (sky rows are cheap, ground rows
     * are not: contiguous bands would not balance). */
    int32_t row_begin;
    int32_t row_count;
    int32_t row_step;
    int32_t reserved;
} rt_render_opts;

#define RT_FLAG_NONE 0
#define RT_FLAG_COUNT_WORK 1 /* also fill the algorithmic-work counters of rt_stats (slower) */
#define RT_FLAG_TILE_SPLIT 2 /* rt_render_multi: split by interleaved scanlines instead of by samples:
                                no accumulator exchange, and the image is bit-identical to the
                                single-GPU render */

typedef struct rt_stats {
    uint64_t samples;      /* (pixel, sample) paths, camera.go:256-260                   */
    uint64_t rays;         /* path segments = world.Hit calls from GetColor, ray.go:37   */
    uint64_t box_tests;    /* only with RT_FLAG_COUNT_WORK                               */
    uint64_t sphere_tests; /* only with RT_FLAG_COUNT_WORK                               */
    uint64_t hits;         /* segments that hit something                                */
    float ms_render;       /* device time, first kernel .. resolved RGB8 in device memory */
    float ms_total;        /* call wall time including copies                            */
    uint32_t kernel_launches;
    uint32_t megakernel_launches; /* render_kernel launches among kernel_launches            */
    float ms_megakernel;          /* summed device time of the render_kernel launches alone   */
    uint32_t reserved;
    uint64_t survivors;    /* paths that outlived their first segment (two-stage mode: queue entries)  */
    uint64_t work_bytes;   /* HBM bytes the call's work buffers carry by construction: a 16-byte radiance
                              record written and read per path, a 48-byte queue entry written and read per
                              survivor, the accumulator read and written once per pass                    */
} rt_stats;

typedef struct rt_scene rt_scene; /* opaque */

/* Thread-local message of the last failure on this thread. */
const char *rt_last_error(void);
int rt_abi_version(void);
/* Number of usable sm_100 devices (0 when there is none; never negative). */
int rt_device_count(void);

/* Replaces NewWorld / World.Add / NewSphere / NewQuad / Box / NewBVHFromWorld (hittables.go:44-53,
 * 85-94, 149-165, 200-216, bvh.go:138-185): validates, derives the quads' w / normal / D, builds the
 * device BVH on the host (binned SAH; topology is free, closest-hit semantics are World.Hit's,
 * hittables.go:55-72), flattens it to 32-byte nodes in depth-first order, and uploads nodes / spheres /
 * quads / materials / texels / Perlin tables once. */
int rt_scene_create(const rt_scene_desc *desc, int device, rt_scene **out);
void rt_scene_destroy(rt_scene *scene);
/* Frees the per-device work buffers the library caches across handles (per-pass radiance buffer,
 * accumulator, RGB8 staging); device < 0 = every device.  They are re-allocated on demand. */
void rt_workspace_release(int device);
/* Run this handle's kernels and copies on the caller's CUDA stream (a cudaStream_t; NULL restores
 * the handle's own stream).  Lets a host framework order the render with its own work (e.g. an
 * NCCL reduce of the accumulators) and time it with events on that stream. */
int rt_scene_set_stream(rt_scene *scene, void *cuda_stream);

/* Replaces the compute half of Camera.Render (camera.go:198-222): GetPixelColor for every pixel
 * (camera.go:254-263), gamma, clamp and quantise (vec3.go:141-166).  rgb_out receives
 * width*height*3 bytes, row-major top to bottom, ready for the "%d %d %d" PPM lines
 * (camera.go:183-188, 242).  accum_out (nullable) receives width*height*3 float32 sample SUMS
 * (before the 1/spp scale of camera.go:261).  Host pointers. */
int rt_render(rt_scene *scene, const rt_camera *camera, const rt_render_opts *opts,
              uint8_t *rgb_out, float *accum_out, rt_stats *stats);

/* The same render on several GPUs of the box from ONE call (a single host process such as the Go
 * program): sample-split.  Device k of `devices` renders its share of the samples on
 * its own host thread; the FP32 accumulators are peer-copied to devices[0], added in device order
 * and resolved there.  Takes the scene description (a scene handle lives on one device).  Every
 * (pixel, sample) keeps its single-GPU Philox stream; only the FP32 summation order differs.
 * With RT_FLAG_TILE_SPLIT the devices take interleaved scanlines (device k: rows k, k+n, ...) at all
 * samples instead and resolve them themselves: no exchange, image bit-identical to rt_render. */
int rt_render_multi(const rt_scene_desc *desc, const rt_camera *camera, const rt_render_opts *opts,
                    const int32_t *devices, int32_t n_devices, uint8_t *rgb_out, float *accum_out,
                    rt_stats *stats);

/* Sample-split building blocks: accumulate `opts->sample_count` samples per pixel,
 * global sample indices [sample_offset, sample_offset+sample_count), into a DEVICE buffer of
 * width*height*3 float32 sums (overwritten, not added to).  The caller reduces the buffers of
 * several GPUs (NCCL) and calls rt_resolve_device on the root. */
int rt_render_accum_device(rt_scene *scene, const rt_camera *camera, const rt_render_opts *opts,
                           float *d_accum, rt_stats *stats);
/* camera.go:261 + camera.go:212-214: scale by 1/total_spp, sqrt, clamp, *255.999, truncate.
 * d_accum: device sums; rgb_out: HOST buffer of width*height*3 bytes.  Runs on cuda_stream (a
 * cudaStream_t, NULL = the default stream) and returns after the copy has completed. */
int rt_resolve_device(const float *d_accum, int32_t width, int32_t height, int32_t total_spp,
                      int32_t device, void *cuda_stream, uint8_t *rgb_out);

/* Parity hook with World.Hit semantics (hittables.go:55-72) on an arbitrary ray batch:
 * origins/dirs are n*3 float32 (host), interval (tmin, tmax) is open (bvh.go:18-20).
 * id_out[i] = object ID of the closest hittable (for a sphere-only scene: its index in
 * rt_scene_desc.spheres) or -1; t_out[i] = its t (unspecified when id is -1). */
int rt_trace(rt_scene *scene, const float *origins, const float *dirs, int64_t n, float tmin,
             float tmax, int32_t *id_out, float *t_out);

/* Parity hook for Camera.GetRay (camera.go:265-299): the device ray generator's rays for pixels
 * [pixel_begin, pixel_begin+n_pixels) x samples [sample_offset, sample_offset+sample_count),
 * written sample-minor as (n_pixels*sample_count)*3 float32 each (host). */
int rt_primary_rays(const rt_camera *camera, const rt_render_opts *opts, int64_t pixel_begin,
                    int64_t n_pixels, float *origins_out, float *dirs_out);

/* Camera.init (camera.go:128-166) on the host. */
int rt_camera_from_options(const rt_camera_options *o, rt_camera *out);

/* Introspection of the flattened device BVH (tests and the work-counting oracle).
 * Node i is 8 x 32-bit words: min.xyz, ref, max.xyz, 0 — its own box and what it contains.
 * ref of an inner node = index of the first of its two children (children are nodes ref and
 * ref+1, siblings adjacent, parents before children: depth-first order); ref of a leaf =
 * 0x80000000 | first_slot << 3 | (count-1), 1..8 spheres (0xC0000000 | ... for a leaf of quads, which
 * have their own slot array).  root_ref describes the root (whose box is not stored); 0xFFFFFFFF
 * for an empty scene.  slot_ids maps a sphere slot to its object ID. */
typedef struct rt_bvh_info {
    uint64_t n_nodes;
    uint64_t n_slots;
    uint32_t max_depth;
    uint32_t in_shared_memory; /* 1 when the scene fits the per-CTA shared-memory staging path */
    float box_pad_min, box_pad_max;
    uint32_t root_ref;
    uint32_t built_on_device; /* 1: the tree was built on the GPU (large sphere-only scenes; Morton-order linear BVH with
                                 leaf collapse) — nodes_out then holds the boxes as the kernels test them; 0: binned SAH on the host */
} rt_bvh_info;
int rt_scene_bvh_info(const rt_scene *scene, rt_bvh_info *out);
int rt_scene_bvh_copy(const rt_scene *scene, uint32_t *nodes_out /* n_nodes*8 */,
                      int32_t *slot_ids_out /* n_slots */);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
