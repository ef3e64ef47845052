// rtgo.hpp — C++ host-side mirror of raytracer-go's scene-description API (package `internal`)
// over the C ABI of librt_b200.so.  The reference is Go and this image has no Go toolchain, so the
// host side above the C ABI is written in C++ with the reference's names, argument meaning and
// error behaviour (camera.go:54-126, 180; hittables.go:39-53, 78-94; materials.go:19-31, 44-58,
// 77-89, 121-173, 297-309; bvh.go:138-140).  A scene written against the reference reads the same:
//
//   auto camera = NewCamera(16.0f / 9.0f, 400, {WithSamplesPerPixel(500), WithMaxRayDepth(50), ...});
//   auto world = NewWorld();
//   world->Add(NewSphere(NewVec3(0, -1000, 0), 1000, NewLambertian(NewCheckered(0.32f, even, odd))));
//   auto tree = NewBVHFromWorld(world);
//   std::string err = camera->Render(tree, out);     // "" on success (Go's nil error)
//
// Only include/rt_b200.h is used: no CUDA headers, no torch.  Nothing here computes colours.
#ifndef RTGO_HPP
#define RTGO_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"

namespace rtgo {

struct Vec3 {
    float X, Y, Z;
};
inline Vec3 NewVec3(float x, float y, float z) { return Vec3{x, y, z}; } // vec3.go:15
inline Vec3 NewVec3Zero() { return Vec3{0, 0, 0}; }                       // vec3.go:23

// ---- textures (materials.go:121-193) ----
struct Texture {
    virtual ~Texture() {}
};
struct SolidColor : Texture {
    Vec3 albedo;
};
struct Checkered : Texture {
    float scale;
    Vec3 even, odd;
};
struct ImageTexture : Texture {
    int w = 0, h = 0;
    std::vector<uint16_t> rgb16; // img.At(i,j).RGBA() r,g,b, row-major
    Vec3 oob{0, 34678.0f * (float)(1.0 / 65535.0), 0}; // zero colour of *image.YCbCr (JPEG)
};
// materials.go:280-295 with Perlin's tables (materials.go:195-216); the reference fills them from a
// clock-seeded *rand.Rand (main.go:120-123), here from a seed.
struct NoiseTexture : Texture {
    rt_perlin perlin;
    float scale;
};
using TexturePtr = std::shared_ptr<Texture>;
inline TexturePtr NewNoiseTexture(uint64_t seed, float scale) {
    auto t = std::make_shared<NoiseTexture>();
    t->scale = scale;
    uint64_t st = seed;
    auto next = [&st]() {
        uint64_t z = (st += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    for (int k = 0; k < 256; k++) // NewVec3RandRange32(-1, 1), materials.go:206
        for (int c = 0; c < 3; c++) t->perlin.vec[k][c] = -1.0f + (float)(next() >> 40) * (1.0f / 16777216.0f) * 2.0f;
    uint8_t *perms[3] = {t->perlin.perm_x, t->perlin.perm_y, t->perlin.perm_z};
    for (auto p : perms) { // Permute(GetNums(256)), materials.go:264-278: target := rand.Intn(i)
        for (int i = 0; i < 256; i++) p[i] = (uint8_t)i;
        for (int i = 255; i > 0; i--) std::swap(p[i], p[next() % (uint64_t)i]);
    }
    return t;
}
inline TexturePtr NewSolidColor(float x, float y, float z) {
    auto t = std::make_shared<SolidColor>();
    t->albedo = NewVec3(x, y, z);
    return t;
}
inline TexturePtr NewCheckered(float scale, Vec3 even, Vec3 odd) {
    auto t = std::make_shared<Checkered>();
    t->scale = scale, t->even = even, t->odd = odd;
    return t;
}
inline TexturePtr NewImageTexture(int w, int h, std::vector<uint16_t> rgb16) {
    auto t = std::make_shared<ImageTexture>();
    t->w = w, t->h = h, t->rgb16 = std::move(rgb16);
    return t;
}

// ---- materials (materials.go:19-119, 297-313) ----
struct Material {
    virtual ~Material() {}
};
struct Lambertian : Material {
    TexturePtr albedo;
};
struct Metal : Material {
    Vec3 albedo;
    float fuzz;
};
struct Dielectric : Material {
    float refractiveIndex;
};
struct DiffuseLight : Material {
    TexturePtr emit;
};
using MaterialPtr = std::shared_ptr<Material>;
inline MaterialPtr NewLambertian(TexturePtr albedo) {
    auto m = std::make_shared<Lambertian>();
    m->albedo = std::move(albedo);
    return m;
}
inline MaterialPtr NewMetal(Vec3 albedo, float fuzz) {
    auto m = std::make_shared<Metal>();
    m->albedo = albedo, m->fuzz = fuzz;
    return m;
}
inline MaterialPtr NewDielectric(float ir) {
    auto m = std::make_shared<Dielectric>();
    m->refractiveIndex = ir;
    return m;
}
inline MaterialPtr NewDiffuseLight(TexturePtr emit) {
    auto m = std::make_shared<DiffuseLight>();
    m->emit = std::move(emit);
    return m;
}

// ---- hittables (hittables.go:39-94, 138-216, bvh.go:132-140) ----
struct HittableObj {
    enum Kind { SPHERE, QUAD } kind;
    Vec3 a, b, c; // sphere: a = Center; quad: a = Q, b = u, c = v
    float Radius;
    MaterialPtr Mat;
};
using Hittable = std::shared_ptr<HittableObj>;
inline Hittable NewSphere(Vec3 center, float radius, MaterialPtr mat) { // hittables.go:85-94
    return std::make_shared<HittableObj>(HittableObj{HittableObj::SPHERE, center, Vec3{}, Vec3{}, radius, std::move(mat)});
}
inline Hittable NewQuad(Vec3 Q, Vec3 u, Vec3 v, MaterialPtr mat) { // hittables.go:149-165 (derived fields: library)
    return std::make_shared<HittableObj>(HittableObj{HittableObj::QUAD, Q, u, v, 0, std::move(mat)});
}
// hittables.go:200-216: the six quads of an axis-aligned box, in the reference's order
inline std::vector<Hittable> Box(Vec3 a, Vec3 b, const MaterialPtr &mat) {
    Vec3 mn{std::fmin(a.X, b.X), std::fmin(a.Y, b.Y), std::fmin(a.Z, b.Z)};
    Vec3 mx{std::fmax(a.X, b.X), std::fmax(a.Y, b.Y), std::fmax(a.Z, b.Z)};
    Vec3 dx{mx.X - mn.X, 0, 0}, dy{0, mx.Y - mn.Y, 0}, dz{0, 0, mx.Z - mn.Z};
    auto neg = [](Vec3 v) { return Vec3{v.X * -1, v.Y * -1, v.Z * -1}; };
    return {NewQuad(NewVec3(mn.X, mn.Y, mx.Z), dx, dy, mat),      NewQuad(NewVec3(mx.X, mn.Y, mx.Z), neg(dz), dy, mat),
            NewQuad(NewVec3(mx.X, mn.Y, mn.Z), neg(dx), dy, mat), NewQuad(NewVec3(mn.X, mn.Y, mn.Z), dz, dy, mat),
            NewQuad(NewVec3(mn.X, mx.Y, mx.Z), dx, neg(dz), mat), NewQuad(NewVec3(mn.X, mn.Y, mn.Z), dx, dz, mat)};
}
struct World {
    std::vector<Hittable> hittables;
    void Add(Hittable h) { hittables.push_back(std::move(h)); } // hittables.go:48-53: insertion order = object ID
    void Add(const std::vector<Hittable> &hs) {
        for (const auto &h : hs) hittables.push_back(h);
    }
};
inline std::shared_ptr<World> NewWorld() { return std::make_shared<World>(); }
// NewBVHFromWorld (bvh.go:138): the reference builds its random-axis pointer tree here; the device
// BVH is built inside rt_scene_create, so the "tree" keeps the insertion-ordered list.
struct BVH {
    std::vector<Hittable> hittables;
};
inline std::shared_ptr<BVH> NewBVHFromWorld(const std::shared_ptr<World> &w) {
    auto b = std::make_shared<BVH>();
    b->hittables = w->hittables;
    return b;
}

// ---- camera (camera.go:23-126) ----
using CameraOpt = std::function<void(rt_camera_options &)>;
inline float ToRadians(float degrees) { return degrees * (float)(M_PI / 180.0); } // math.go:46-52
inline CameraOpt WithSamplesPerPixel(int n) { return [=](rt_camera_options &o) { o.spp = n; }; }
inline CameraOpt WithMaxRayDepth(int n) { return [=](rt_camera_options &o) { o.max_depth = n; }; }
inline CameraOpt WithFOVDegrees(float f) { return [=](rt_camera_options &o) { o.fov_radians = ToRadians(f); }; }
inline CameraOpt WithLookAt(Vec3 v) { return [=](rt_camera_options &o) { o.look_at[0] = v.X, o.look_at[1] = v.Y, o.look_at[2] = v.Z; }; }
inline CameraOpt WithLookFrom(Vec3 v) { return [=](rt_camera_options &o) { o.look_from[0] = v.X, o.look_from[1] = v.Y, o.look_from[2] = v.Z; }; }
inline CameraOpt WithDefocusAngleDegrees(float d) { return [=](rt_camera_options &o) { o.defocus_angle_radians = ToRadians(d); }; }
inline CameraOpt WithFocusDist(float d) { return [=](rt_camera_options &o) { o.focus_dist = d; }; }
inline CameraOpt WithBackgroundColor(Vec3 c) { return [=](rt_camera_options &o) { o.background[0] = c.X, o.background[1] = c.Y, o.background[2] = c.Z; }; }

// PNG without a compression library: 8-bit truecolour, filter 0, the zlib stream made of stored
// (uncompressed) deflate blocks — a valid PNG every decoder reads, 3 bytes per pixel + 0.01 % framing.
inline void write_png(std::ostream &out, const uint8_t *rgb, int w, int h) {
    static uint32_t table[256];
    static bool have_table = false;
    if (!have_table) {
        for (uint32_t n = 0; n < 256; n++) {
            uint32_t c = n;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[n] = c;
        }
        have_table = true;
    }
    auto be32 = [](std::string &s, uint32_t v) {
        s.push_back((char)(v >> 24)), s.push_back((char)(v >> 16)), s.push_back((char)(v >> 8)), s.push_back((char)v);
    };
    auto chunk = [&](const char *tag, const std::string &data) {
        std::string c;
        be32(c, (uint32_t)data.size());
        c.append(tag, 4);
        c += data;
        uint32_t crc = 0xFFFFFFFFu;
        for (size_t i = 4; i < c.size(); i++) crc = table[(crc ^ (uint8_t)c[i]) & 0xFF] ^ (crc >> 8);
        be32(c, crc ^ 0xFFFFFFFFu);
        out.write(c.data(), (std::streamsize)c.size());
    };
    out.write("\x89PNG\r\n\x1a\n", 8);
    std::string ihdr;
    be32(ihdr, (uint32_t)w), be32(ihdr, (uint32_t)h);
    ihdr += std::string("\x08\x02\x00\x00\x00", 5); // 8 bits, colour type 2 (RGB), deflate, filter 0, no interlace
    chunk("IHDR", ihdr);
    // raw image data: a filter byte (0) in front of every scanline
    std::string raw;
    raw.reserve((size_t)h * ((size_t)w * 3 + 1));
    for (int j = 0; j < h; j++) {
        raw.push_back('\0');
        raw.append(reinterpret_cast<const char *>(rgb) + (size_t)j * w * 3, (size_t)w * 3);
    }
    std::string z("\x78\x01", 2); // zlib header: deflate, 32 K window, no preset dictionary
    uint32_t a = 1, b = 0;         // adler32 of the raw data
    for (size_t pos = 0; pos < raw.size() || pos == 0;) {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        const bool last = pos + n >= raw.size();
        z.push_back(last ? '\x01' : '\x00'); // stored block, BFINAL on the last one
        z.push_back((char)(n & 0xFF)), z.push_back((char)(n >> 8));
        z.push_back((char)(~n & 0xFF)), z.push_back((char)((~n >> 8) & 0xFF));
        z.append(raw, pos, n);
        for (size_t i = pos; i < pos + n; i++) {
            a = (a + (uint8_t)raw[i]) % 65521u;
            b = (b + a) % 65521u;
        }
        pos += n;
        if (last) break;
    }
    be32(z, (b << 16) | a);
    chunk("IDAT", z);
    chunk("IEND", std::string());
}

struct Camera {
    rt_camera_options options{};
    rt_camera c{};
    uint64_t seed = 0xC0FFEE; // no counterpart in the reference (clock-seeded, camera.go:170)
    int device = 0;
    rt_stats last_stats{};

    // camera.go:180: returns "" for Go's nil error, the message otherwise.  binary = false writes the
    // reference's P3 text; binary = true writes P6 (the reference's TODO at camera.go:196).
    std::string Render(const std::shared_ptr<BVH> &world, std::ostream &writer, bool binary = false, bool png = false) {
        // flatten: materials / textures de-duplicated by pointer, spheres in insertion order
        std::vector<rt_sphere> spheres;
        std::vector<rt_quad> quads;
        std::vector<uint32_t> sphere_ids, quad_ids;
        std::vector<rt_material> materials;
        std::vector<rt_texture> textures;
        std::vector<rt_image> images;
        std::vector<rt_perlin> perlins;
        std::map<const Material *, uint32_t> mat_index;
        std::map<const Texture *, uint32_t> tex_index;
        auto tex_id = [&](const TexturePtr &t, uint32_t *out) -> bool {
            auto it = tex_index.find(t.get());
            if (it != tex_index.end()) return *out = it->second, true;
            rt_texture r{};
            if (auto s = dynamic_cast<const SolidColor *>(t.get())) {
                r.kind = RT_TEX_SOLID, r.a[0] = s->albedo.X, r.a[1] = s->albedo.Y, r.a[2] = s->albedo.Z;
            } else if (auto c = dynamic_cast<const Checkered *>(t.get())) {
                r.kind = RT_TEX_CHECKER, r.scale = c->scale;
                r.a[0] = c->even.X, r.a[1] = c->even.Y, r.a[2] = c->even.Z;
                r.b[0] = c->odd.X, r.b[1] = c->odd.Y, r.b[2] = c->odd.Z;
            } else if (auto im = dynamic_cast<const ImageTexture *>(t.get())) {
                r.kind = RT_TEX_IMAGE, r.image = (uint32_t)images.size();
                r.oob[0] = im->oob.X, r.oob[1] = im->oob.Y, r.oob[2] = im->oob.Z;
                images.push_back(rt_image{im->w, im->h, im->rgb16.data()});
            } else if (auto nt = dynamic_cast<const NoiseTexture *>(t.get())) {
                r.kind = RT_TEX_NOISE, r.scale = nt->scale, r.image = (uint32_t)perlins.size();
                perlins.push_back(nt->perlin);
            } else {
                return false;
            }
            textures.push_back(r);
            return *out = tex_index[t.get()] = (uint32_t)textures.size() - 1, true;
        };
        for (const auto &h : world->hittables) {
            const Material *m = h->Mat.get();
            auto it = mat_index.find(m);
            uint32_t mi;
            if (it != mat_index.end()) {
                mi = it->second;
            } else {
                rt_material r{};
                if (auto l = dynamic_cast<const Lambertian *>(m)) {
                    r.kind = RT_MAT_LAMBERTIAN;
                    if (!tex_id(l->albedo, &r.texture)) return "b200: texture is outside the accelerated path";
                } else if (auto me = dynamic_cast<const Metal *>(m)) {
                    r.kind = RT_MAT_METAL, r.fuzz = me->fuzz;
                    r.albedo[0] = me->albedo.X, r.albedo[1] = me->albedo.Y, r.albedo[2] = me->albedo.Z;
                } else if (auto d = dynamic_cast<const Dielectric *>(m)) {
                    r.kind = RT_MAT_DIELECTRIC, r.ior = d->refractiveIndex;
                } else if (auto dl = dynamic_cast<const DiffuseLight *>(m)) {
                    r.kind = RT_MAT_DIFFUSE_LIGHT;
                    if (!tex_id(dl->emit, &r.texture)) return "b200: texture is outside the accelerated path";
                } else {
                    return "b200: material is outside the accelerated path";
                }
                materials.push_back(r);
                mi = mat_index[m] = (uint32_t)materials.size() - 1;
            }
            const uint32_t object_id = (uint32_t)(spheres.size() + quads.size());
            if (h->kind == HittableObj::SPHERE) {
                spheres.push_back(rt_sphere{h->a.X, h->a.Y, h->a.Z, h->Radius, mi});
                sphere_ids.push_back(object_id);
            } else {
                quads.push_back(rt_quad{{h->a.X, h->a.Y, h->a.Z}, {h->b.X, h->b.Y, h->b.Z}, {h->c.X, h->c.Y, h->c.Z}, mi});
                quad_ids.push_back(object_id);
            }
        }
        rt_scene_desc desc{};
        desc.abi_version = RT_B200_ABI_VERSION;
        desc.spheres = spheres.data(), desc.n_spheres = spheres.size();
        desc.materials = materials.data(), desc.n_materials = (uint32_t)materials.size();
        desc.textures = textures.data(), desc.n_textures = (uint32_t)textures.size();
        desc.images = images.data(), desc.n_images = (uint32_t)images.size();
        desc.perlins = perlins.data(), desc.n_perlins = (uint32_t)perlins.size();
        desc.quads = quads.data(), desc.n_quads = quads.size();
        desc.sphere_ids = sphere_ids.data(), desc.quad_ids = quad_ids.data();
        rt_scene *scene = nullptr;
        if (rt_scene_create(&desc, device, &scene) != RT_OK) return rt_last_error();
        const int w = c.width, h = c.height;
        std::vector<uint8_t> rgb((size_t)w * h * 3);
        rt_render_opts ro{};
        ro.seed = seed, ro.device = device;
        int rc = rt_render(scene, &c, &ro, rgb.data(), nullptr, &last_stats);
        std::string err = rc == RT_OK ? "" : rt_last_error();
        rt_scene_destroy(scene);
        if (!err.empty()) return err;
        if (png) {
            write_png(writer, rgb.data(), w, h);
            return writer ? "" : "write failed";
        }
        if (binary) {
            writer << "P6\n" << w << " " << h << "\n255\n";
            writer.write(reinterpret_cast<const char *>(rgb.data()), (std::streamsize)rgb.size());
            return writer ? "" : "write failed";
        }
        // camera.go:183-188 header, then one "R G B" line per pixel in chunks of 5000 (camera.go:225, 242)
        writer << "P3\n" << w << " " << h << "\n255\n";
        std::string chunk;
        for (size_t p = 0; p < (size_t)w * h; p++) {
            chunk += std::to_string(rgb[3 * p]) + " " + std::to_string(rgb[3 * p + 1]) + " " + std::to_string(rgb[3 * p + 2]) + "\n";
            if ((p + 1) % 5000 == 0 || p + 1 == (size_t)w * h) {
                writer << chunk;
                chunk.clear();
                if (!writer) return "write failed";
            }
        }
        return "";
    }
};

// camera.go:104-126 with its defaults (camera.go:105-117); Camera.init runs in rt_camera_from_options.
inline std::shared_ptr<Camera> NewCamera(float aspectRatio, int imageWidth, const std::vector<CameraOpt> &opts) {
    auto cam = std::make_shared<Camera>();
    rt_camera_options &o = cam->options;
    o.aspect_ratio = aspectRatio, o.image_width = imageWidth;
    o.fov_radians = (float)(M_PI / 2);
    o.spp = 100, o.max_depth = 50, o.focus_dist = 10, o.defocus_angle_radians = 0;
    o.look_at[0] = o.look_at[1] = o.look_at[2] = 0;
    o.look_from[0] = 0, o.look_from[1] = 0, o.look_from[2] = -1;
    o.vup[0] = 0, o.vup[1] = 1, o.vup[2] = 0;
    o.background[0] = o.background[1] = o.background[2] = 0;
    for (const auto &fn : opts) fn(o);
    rt_camera_from_options(&o, &cam->c);
    return cam;
}

} // namespace rtgo
#endif
