// rt_shade.h — ray generation, hit completion, materials, textures and pixel resolve, in the
// reference's float32 operation order.  __host__ __device__ for the same reason as rt_math.h.
//
//   camera.go:265-299   GetRay / sampleUnitSquare        -> generate_ray
//   hittables.go:118-128, 22-37  point, normal, front    -> complete_hit
//   hittables.go:122-126 sphere UV (only image textures) -> sphere_uv
//   materials.go:33-42 / 60-75 / 91-119 / 301-313        -> shade_hit
//   materials.go:127-137 / 155-157 / 175-193             -> texture_value
//   camera.go:261 + vec3.go:145-166 + 141-143            -> resolve_pixel
#ifndef RT_SHADE_H
#define RT_SHADE_H

#include "rt_math.h"
#include "rt_rng.h"
#include "rt_trace.h"
#include "../../include/rt_b200.h"

// Device material record = 2 x F4 (texture folded into the material at upload):
//   m0 = (c0.r, c0.g, c0.b, p0)   m1 = (c1.r, c1.g, c1.b, bits(code))
//   code = material kind | texture kind << 4 | image index << 8
//   Lambertian/DiffuseLight + solid  : c0 = colour
//                          + checker : c0 = even, c1 = odd, p0 = 1/scale (materials.go:128)
//                          + image   : c0 = out-of-bounds colour
//   Metal      : c0 = albedo, p0 = fuzz
//   Dielectric : p0 = ior, c1.r = 1/ior (materials.go:94)
#define RT_CODE(mat, tex, img) ((uint32_t)(mat) | ((uint32_t)(tex) << 4) | ((uint32_t)(img) << 8))
#define RT_CODE_MAT(c) ((c) & 15u)
#define RT_CODE_TEX(c) (((c) >> 4) & 15u)
#define RT_CODE_IMG(c) ((c) >> 8)

struct DevImage {
    const uint16_t *texels; // 4 x uint16 per texel (r, g, b, 0), row-major
    int32_t w, h;
};

// Perlin's tables on the device (materials.go:195-200): gradients as F4, the three permutations.
struct DevPerlin {
    F4 vec[256];
    uint8_t perm_x[256], perm_y[256], perm_z[256];
};

// What Texture.GetTexture may have to read besides the material record.
struct DevTex {
    const DevImage *images;
    const DevPerlin *perlins;
};

struct DevCamera {
    V3 center, pixel00, du, dv, disk_u, disk_v, background;
    int32_t width, height, max_depth;
    int32_t defocus; // defocusAngleRadians > 0, camera.go:279
};

RT_HD DevCamera make_dev_camera(const rt_camera &c) {
    DevCamera d;
    d.center = v3(c.center[0], c.center[1], c.center[2]);
    d.pixel00 = v3(c.pixel00[0], c.pixel00[1], c.pixel00[2]);
    d.du = v3(c.pixel_du[0], c.pixel_du[1], c.pixel_du[2]);
    d.dv = v3(c.pixel_dv[0], c.pixel_dv[1], c.pixel_dv[2]);
    d.disk_u = v3(c.defocus_u[0], c.defocus_u[1], c.defocus_u[2]);
    d.disk_v = v3(c.defocus_v[0], c.defocus_v[1], c.defocus_v[2]);
    d.background = v3(c.background[0], c.background[1], c.background[2]);
    d.width = c.width, d.height = c.height, d.max_depth = c.max_depth;
    d.defocus = c.defocus_angle > 0 ? 1 : 0;
    return d;
}

// camera.go:265-299.  Draw order (rt_rng.h): one block = (dx, dy, disk.x, disk.y); while the disk
// pair fails x^2+y^2 < 1 (vec3.go:203-210; the disk sample is always drawn, camera.go:277) another
// block supplies two more candidate pairs.
RT_HD void generate_ray(const DevCamera &c, PathRng &rng, int i, int j, V3 &origin, V3 &dir) {
    V3 du_off = c.du * (float)i;
    V3 dv_off = c.dv * (float)j;
    V3 pc = c.pixel00;
    pc = pc + du_off;
    pc = pc + dv_off;
    RngBlock b = rng.next();
    float dx = rand_centered(b.w0); // -0.5 + Float32(), camera.go:290
    float dy = rand_centered(b.w1);
    pc = pc + (c.du * dx + c.dv * dy);
    float sx = rand_pm1(b.w2), sy = rand_pm1(b.w3);
    while (!(sx * sx + sy * sy + 0.0f * 0.0f < 1.0f)) {
        b = rng.next();
        sx = rand_pm1(b.w0), sy = rand_pm1(b.w1);
        if (sx * sx + sy * sy + 0.0f * 0.0f < 1.0f) break;
        sx = rand_pm1(b.w2), sy = rand_pm1(b.w3);
    }
    origin = c.center;
    if (c.defocus) origin = c.center + (c.disk_u * sx + c.disk_v * sy);
    dir = pc - origin;
}

// The beam of pixel (i, j): every ray generate_ray can return for it, whatever the sample (rt_trace.h: Beam).  Origins
// lie in the box around the camera centre that holds the defocus parallelogram (|sx|, |sy| <= 1), the points aimed
// at in the pixel's footprint (|dx|, |dy| <= 0.5) around the centre formed exactly as generate_ray forms it; both are
// widened by RT_BEAM_EPS of the magnitudes involved — the float32 roundings of generate_ray's sums are ~1e-7 of them.
RT_HD Beam pixel_beam(const DevCamera &c, int i, int j) {
    V3 pc = c.pixel00;
    pc = pc + c.du * (float)i;
    pc = pc + c.dv * (float)j;
    const V3 hw = v3(0.5f * (fabsf(c.du.x) + fabsf(c.dv.x)), 0.5f * (fabsf(c.du.y) + fabsf(c.dv.y)), 0.5f * (fabsf(c.du.z) + fabsf(c.dv.z)));
    const V3 ow = c.defocus ? v3(fabsf(c.disk_u.x) + fabsf(c.disk_v.x), fabsf(c.disk_u.y) + fabsf(c.disk_v.y), fabsf(c.disk_u.z) + fabsf(c.disk_v.z))
                            : v3(0, 0, 0);
    const V3 m = v3(RT_BEAM_EPS * (fabsf(pc.x) + fabsf(c.center.x) + hw.x + ow.x), RT_BEAM_EPS * (fabsf(pc.y) + fabsf(c.center.y) + hw.y + ow.y),
                    RT_BEAM_EPS * (fabsf(pc.z) + fabsf(c.center.z) + hw.z + ow.z));
    Beam b;
    b.olo = c.center - ow - m, b.ohi = c.center + ow + m;
    b.dlo = (pc - hw - m) - b.ohi, b.dhi = (pc + hw + m) - b.olo;
    return b;
}

struct HitInfo {
    V3 point, normal;
    bool front;
    bool is_quad;  // u, v below are valid (a quad's alpha, beta); a sphere's UV is computed on demand
    float u, v;
};

// hittables.go:118-120 and NewHitInfo (hittables.go:22-37)
RT_HD void complete_hit(const F4 &s, V3 o, V3 d, float t, HitInfo &hi) {
    V3 point = d * t + o;                                 // ray.go:25-30
    V3 norm = unit((point - v3(s.x, s.y, s.z)) * s.w);    // hittables.go:119-120
    bool front = dot(d, norm) < 0;
    if (!front) norm = norm * -1.0f;
    hi.point = point, hi.normal = norm, hi.front = front;
    hi.is_quad = false, hi.u = 0, hi.v = 0;
}

// Quad.Hit's HitInfo (hittables.go:180-190): point, alpha/beta as (u, v), the quad's normal flipped
// against the ray.  `q` is the 5 x F4 device record (rt_trace.h).
RT_HD void complete_hit_quad(const F4 *__restrict__ q, V3 o, V3 d, float t, HitInfo &hi) {
    const V3 p = d * t + o;
    const V3 ph = p - v3(q[0].x, q[0].y, q[0].z);
    const V3 w = v3(q[3].x, q[3].y, q[3].z);
    hi.u = dot(w, cross(ph, v3(q[2].x, q[2].y, q[2].z)));
    hi.v = dot(w, cross(v3(q[1].x, q[1].y, q[1].z), ph));
    V3 norm = v3(q[4].x, q[4].y, q[4].z);
    const bool front = dot(d, norm) < 0;
    if (!front) norm = norm * -1.0f;
    hi.point = p, hi.normal = norm, hi.front = front, hi.is_quad = true;
}

// hittables.go:122-126.  `outward` is the normal BEFORE the front-face flip.
RT_HD void sphere_uv(V3 outward, float &u, float &v) {
    const float pi32 = 3.14159265358979323846f; // math.go:48
    float theta = (float)acos(-(double)outward.y);
    float phi = (float)(atan2(-(double)outward.z, (double)outward.x) + 3.14159265358979323846);
    u = div32(phi + div32(5 * pi32, 12.0f), 2 * pi32);
    v = div32(theta, pi32);
}

// materials.go:175-193
RT_HD V3 image_texture(const DevImage &im, V3 oob, float u, float v) {
    if (im.h <= 0) return v3(0, 1, 1);
    u = clamp01(u);
    v = 1 - clamp01(v);
    float fi = u * (float)im.w;
    float fj = v * (float)im.h;
    int i = (int)fi, j = (int)fj;
    if (i < 0 || i >= im.w || j < 0 || j >= im.h) return oob; // image.At outside Bounds()
    RT_DBG((size_t)j * (size_t)im.w + (size_t)i < (size_t)im.w * (size_t)im.h, RT_DBG_TEXEL);
    const uint16_t *px = im.texels + ((size_t)j * (size_t)im.w + (size_t)i) * 4;
#if defined(__CUDA_ARCH__)
    const ushort4 t = *reinterpret_cast<const ushort4 *>(px);
    const uint16_t r = t.x, g = t.y, b = t.z;
#else
    const uint16_t r = px[0], g = px[1], b = px[2];
#endif
    const float col_scale = (float)(1.0 / 65535.0);
    return v3((float)r * col_scale, (float)g * col_scale, (float)b * col_scale);
}

// math.go:58-60, 78-92: Lerp, BiLinearLerp, TriLinearLerp
RT_HD float lerp1(float t, float x, float y) { return x * (1 - t) + y * t; }
RT_HD float bilerp(float tx, float ty, float c00, float c10, float c01, float c11) {
    const float a = lerp1(tx, c00, c10);
    const float b = lerp1(tx, c01, c11);
    return lerp1(ty, a, b);
}
// materials.go:218-220
RT_HD float smoothstep(float t) { return t * t * (3 - 2 * t); }
// Perlin.Noise, materials.go:223-249
RT_HD float perlin_noise(const DevPerlin &per, V3 p) {
    const float xi = floorf(p.x), yi = floorf(p.y), zi = floorf(p.z);
    const float tx = p.x - xi, ty = p.y - yi, tz = p.z - zi;
    const int rx0 = (int)((long long)xi & 255), rx1 = (rx0 + 1) & 255;
    const int ry0 = (int)((long long)yi & 255), ry1 = (ry0 + 1) & 255;
    const int rz0 = (int)((long long)zi & 255), rz1 = (rz0 + 1) & 255;
    const uint32_t x0 = per.perm_x[rx0], x1 = per.perm_x[rx1], y0 = per.perm_y[ry0], y1 = per.perm_y[ry1];
    const uint32_t z0 = per.perm_z[rz0], z1 = per.perm_z[rz1];
#define RT_GRAD(ix, iy, iz, ax, ay, az) dot(v3(per.vec[ix ^ iy ^ iz].x, per.vec[ix ^ iy ^ iz].y, per.vec[ix ^ iy ^ iz].z), v3(ax, ay, az))
    const float c000 = RT_GRAD(x0, y0, z0, tx, ty, tz);
    const float c001 = RT_GRAD(x0, y0, z1, tx, ty, tz - 1);
    const float c010 = RT_GRAD(x0, y1, z0, tx, ty - 1, tz);
    const float c011 = RT_GRAD(x0, y1, z1, tx, ty - 1, tz - 1);
    const float c100 = RT_GRAD(x1, y0, z0, tx - 1, ty, tz);
    const float c101 = RT_GRAD(x1, y0, z1, tx - 1, ty, tz - 1);
    const float c110 = RT_GRAD(x1, y1, z0, tx - 1, ty - 1, tz);
    const float c111 = RT_GRAD(x1, y1, z1, tx - 1, ty - 1, tz - 1);
#undef RT_GRAD
    const float sx = smoothstep(tx), sy = smoothstep(ty), sz = smoothstep(tz);
    const float e = bilerp(sx, sy, c000, c100, c010, c110);
    const float f = bilerp(sx, sy, c001, c101, c011, c111);
    return lerp1(sz, e, f);
}
// Perlin.Turb, materials.go:251-262
RT_HD float perlin_turb(const DevPerlin &per, V3 p, int depth) {
    float sum = 0, weight = 1.0f;
    for (int i = 0; i < depth; i++) {
        sum += weight * perlin_noise(per, p);
        weight *= 0.5f;
        p = p * 2.0f;
    }
    return fabsf(sum);
}
// NoiseTexture.GetTexture, materials.go:285-288 (sin in f64 as math.Sin)
RT_HD V3 noise_texture(const DevPerlin &per, float scale, V3 point) {
    point = point * scale;
    const float s = 0.5f * (1 + (float)sin((double)(point.z + 10 * perlin_turb(per, point, 7))));
    return v3(1, 1, 1) * s;
}

// Texture.GetTexture for the texture folded into material record (m0, m1).
RT_HD V3 texture_value(const F4 &m0, const F4 &m1, uint32_t code, DevTex tex_, const HitInfo &hi) {
    const V3 point = hi.point;
    const uint32_t tex = RT_CODE_TEX(code);
    if (tex == RT_TEX_CHECKER) { // materials.go:127-137
        const float inv = m0.w;
        int x = (int)floorf(inv * point.x);
        int y = (int)floorf(inv * point.y);
        int z = (int)floorf(inv * point.z);
        if (((x + y + z) & 1) == 0) return v3(m0.x, m0.y, m0.z);
        return v3(m1.x, m1.y, m1.z);
    }
    RT_DBG(tex != RT_TEX_NOISE || RT_CODE_IMG(code) < RT_DBG_B(n_perlins), RT_DBG_TEXTURE);
    RT_DBG(tex != RT_TEX_IMAGE || RT_CODE_IMG(code) < RT_DBG_B(n_images), RT_DBG_TEXTURE);
    if (tex == RT_TEX_NOISE) return noise_texture(tex_.perlins[RT_CODE_IMG(code)], m0.w, point);
    if (tex == RT_TEX_IMAGE) {
        float u = hi.u, v = hi.v;
        if (!hi.is_quad) sphere_uv(hi.front ? hi.normal : hi.normal * -1.0f, u, v); // the outward normal
        return image_texture(tex_.images[RT_CODE_IMG(code)], v3(m0.x, m0.y, m0.z), u, v);
    }
    return v3(m0.x, m0.y, m0.z); // materials.go:155-157
}

// materials.go:115-119
RT_HD float reflectance(float cos_theta, float eta) {
    float r0 = div32(1.0f - eta, 1.0f + eta);
    r0 *= r0;
    double x = 1 - (double)cos_theta;
    double x2 = x * x;
    float p5 = (float)(x2 * x2 * x); // math.Pow(x, 5)
    return r0 + (1 - r0) * p5;
}

// One Emit + Scatter (ray.go:41-50).  Returns false when the path ends (no scatter); `emitted`
// is always written.  On scatter, (o, d) become the scattered ray and `atten` its attenuation.
// The work shared by several materials (the unit-sphere sample of Lambertian and Metal, Unit(dir)
// of Metal and Dielectric) is hoisted so that lanes with different materials run it together.
RT_HD bool shade_surface(const F4 &m0, const F4 &m1, DevTex tex, const HitInfo &hi,
                         PathRng &rng, V3 &o, V3 &d, V3 &atten, V3 &emitted) {
    const uint32_t code = as_uint(m1.w);
    const uint32_t kind = RT_CODE_MAT(code);
    emitted = v3(0, 0, 0);
    V3 ru = v3(0, 0, 0), unit_dir = v3(0, 0, 0);
    if (kind == RT_MAT_LAMBERTIAN || kind == RT_MAT_METAL) ru = rand_unit(rng); // materials.go:34, 64
    if (kind == RT_MAT_METAL || kind == RT_MAT_DIELECTRIC) unit_dir = unit(d);  // materials.go:61, 97
    if (kind == RT_MAT_LAMBERTIAN) { // materials.go:33-42
        V3 dir = hi.normal + ru;
        if (near_zero(dir)) dir = hi.normal;
        atten = texture_value(m0, m1, code, tex, hi);
        o = hi.point, d = dir;
        return true;
    }
    if (kind == RT_MAT_METAL) { // materials.go:60-75
        V3 reflected = reflect(unit_dir, hi.normal);
        V3 fuzz = ru * m0.w;
        V3 scattered = reflected + fuzz;
        if (dot(scattered, hi.normal) > 0) {
            atten = v3(m0.x, m0.y, m0.z);
            o = hi.point, d = scattered;
            return true;
        }
        return false;
    }
    if (kind == RT_MAT_DIELECTRIC) { // materials.go:91-113
        const float eta = hi.front ? m1.x : m0.w;
        float cos_theta = fminf(dot(unit_dir * -1.0f, hi.normal), 1.0f);
        float sin_theta = (float)sqrt(1 - (double)(cos_theta * cos_theta));
        bool cannot_refract = sin_theta * eta > 1.0f;
        V3 direction;
        // `||` short-circuits (materials.go:103): the uniform is drawn only if refraction is possible
        if (cannot_refract || reflectance(cos_theta, eta) > rng.next().u0)
            direction = reflect(unit_dir, hi.normal);
        else
            direction = refract(unit_dir, hi.normal, eta);
        atten = v3(1, 1, 1);
        o = hi.point, d = direction;
        return true;
    }
    // DiffuseLight: emits its texture, never scatters (materials.go:301-313)
    emitted = texture_value(m0, m1, code, tex, hi);
    return false;
}

// Sphere hit (hittables.go:118-128) then Emit + Scatter.
RT_HD bool shade_hit(const F4 &m0, const F4 &m1, DevTex tex, const F4 &sphere, float t,
                     PathRng &rng, V3 &o, V3 &d, V3 &atten, V3 &emitted) {
    HitInfo hi;
    complete_hit(sphere, o, d, t, hi);
    return shade_surface(m0, m1, tex, hi, rng, o, d, atten, emitted);
}

// Quad hit (hittables.go:180-190) then Emit + Scatter.
RT_HD bool shade_hit_quad(const F4 &m0, const F4 &m1, DevTex tex, const F4 *__restrict__ quad, float t,
                          PathRng &rng, V3 &o, V3 &d, V3 &atten, V3 &emitted) {
    HitInfo hi;
    complete_hit_quad(quad, o, d, t, hi);
    return shade_surface(m0, m1, tex, hi, rng, o, d, atten, emitted);
}

// camera.go:261 (sum * (1/spp)), vec3.go:162-166 (sqrt), 145-152 (clamp, *255.999), 141-143 (int())
RT_HD void resolve_pixel(V3 sum, float inv_spp, uint8_t *rgb) {
    V3 mean = sum * inv_spp;
    float ch[3] = {mean.x, mean.y, mean.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float g = sqrt32(ch[k]);
        g = clamp01(g);
        g *= 255.999f;
        rgb[k] = (g != g) ? (uint8_t)0 : (uint8_t)(int)g;
    }
}

#endif // RT_SHADE_H
