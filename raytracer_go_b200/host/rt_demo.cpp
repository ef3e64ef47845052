// rt_demo.cpp — main.go's randSpheres (main.go:227-289) and cornellBox (main.go:194-225) scenes written
// against the C++ mirror of the reference API and rendered through librt_b200.so; writes a P3 PPM like
// the reference's out/img.ppm.
//   rt_demo [width=400] [spp=500] [out=out/img.ppm] [scene_seed=0x5EED0001]      random spheres
//   rt_demo cornell [width=600] [spp=200] [out=out/img.ppm]                       Cornell box
// (an `out` ending in .png is written as PNG)
// The reference seeds its scene RNG from the clock (main.go:246); here a fixed-seed SplitMix64
// supplies rand.Float32() so runs are reproducible.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "rtgo.hpp"

using namespace rtgo;

struct Rand {
    uint64_t s;
    float Float32() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (float)(z >> 40) * (1.0f / 16777216.0f);
    }
};

// an output path ending in .png selects the PNG writer, anything else the reference's P3 text
static bool ends_with_png(const char *path) {
    const std::string p(path);
    return p.size() >= 4 && p.compare(p.size() - 4, 4, ".png") == 0;
}

static int cornell(int argc, char **argv) { // main.go:194-225
    const int width = argc > 2 ? atoi(argv[2]) : 600;
    const int spp = argc > 3 ? atoi(argv[3]) : 200;
    const char *out = argc > 4 ? argv[4] : "out/img.ppm";
    auto camera = NewCamera(1, width,
                            {WithSamplesPerPixel(spp), WithMaxRayDepth(50), WithLookFrom(NewVec3(278, 278, -800)),
                             WithLookAt(NewVec3(278, 278, 0)), WithFOVDegrees(40), WithDefocusAngleDegrees(0),
                             WithBackgroundColor(NewVec3Zero())});
    auto world = NewWorld();
    auto red = NewLambertian(NewSolidColor(.65f, .05f, .05f));
    auto white = NewLambertian(NewSolidColor(.73f, .73f, .73f));
    auto green = NewLambertian(NewSolidColor(.12f, .45f, .15f));
    auto light = NewDiffuseLight(NewSolidColor(15, 15, 15));
    world->Add(NewQuad(NewVec3(555, 0, 0), NewVec3(0, 555, 0), NewVec3(0, 0, 555), green));
    world->Add(NewQuad(NewVec3(0, 0, 0), NewVec3(0, 555, 0), NewVec3(0, 0, 555), red));
    world->Add(NewQuad(NewVec3(343, 554, 332), NewVec3(-130, 0, 0), NewVec3(0, 0, -105), light));
    world->Add(NewQuad(NewVec3(0, 0, 0), NewVec3(555, 0, 0), NewVec3(0, 0, 555), white));
    world->Add(NewQuad(NewVec3(555, 555, 555), NewVec3(-555, 0, 0), NewVec3(0, 0, -555), white));
    world->Add(NewQuad(NewVec3(0, 0, 555), NewVec3(555, 0, 0), NewVec3(0, 555, 0), white));
    world->Add(Box(NewVec3(130, 0, 65), NewVec3(295, 165, 230), white));
    world->Add(Box(NewVec3(265, 0, 295), NewVec3(430, 330, 460), white));
    std::ofstream f(out, std::ios::binary);
    if (!f) return fprintf(stderr, "cannot open %s\n", out), 2;
    std::string err = camera->Render(NewBVHFromWorld(world), f, false, ends_with_png(out));
    if (!err.empty()) return fprintf(stderr, "render failed: %s\n", err.c_str()), 1;
    const rt_stats &st = camera->last_stats;
    printf("Finished: cornell box %dx%d, %d spp: %.1f Msamples/s device, %.1f Mrays/s\n", camera->c.width,
           camera->c.height, spp, st.samples / (st.ms_render * 1e3), st.rays / (st.ms_render * 1e3));
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "cornell") return cornell(argc, argv);
    const int width = argc > 1 ? atoi(argv[1]) : 400;
    const int spp = argc > 2 ? atoi(argv[2]) : 500;
    const char *out = argc > 3 ? argv[3] : "out/img.ppm";
    Rand rnd{argc > 4 ? strtoull(argv[4], nullptr, 0) : 0x5EED0001ull};
    auto t0 = std::chrono::steady_clock::now();

    auto camera = NewCamera(16.0f / 9.0f, width,
                            {WithSamplesPerPixel(spp), WithMaxRayDepth(50), WithLookFrom(NewVec3(13, 2, 3)),
                             WithLookAt(NewVec3(0, 0, 0)), WithFOVDegrees(20), WithDefocusAngleDegrees(0.6f),
                             WithFocusDist(10), WithBackgroundColor(NewVec3(0.7f, 0.8f, 1))});
    auto world = NewWorld();
    auto checkered = NewCheckered(0.32f, NewVec3(0.2f, 0.3f, 0.1f), NewVec3(0.9f, 0.9f, 0.9f));
    world->Add(NewSphere(NewVec3(0, -1000, 0), 1000, NewLambertian(checkered)));
    for (int i = -11; i < 11; i++) {
        for (int j = -11; j < 11; j++) {
            float matPer = rnd.Float32();
            float cx = (float)i + 0.9f * rnd.Float32();
            float cz = (float)j + 0.9f * rnd.Float32();
            float dx = cx - 4, dy = 0.2f - 0.2f, dz = cz - 0;
            if (std::sqrt(dx * dx + dy * dy + dz * dz) > 0.9f) {
                MaterialPtr m;
                if (matPer < 0.8f) {
                    float a = rnd.Float32(), b = rnd.Float32(), c = rnd.Float32();
                    float d = rnd.Float32(), e = rnd.Float32(), f = rnd.Float32();
                    m = NewLambertian(NewSolidColor(a * d, b * e, c * f));
                } else if (matPer < 0.95f) {
                    float a = 0.5f + rnd.Float32() * 0.5f, b = 0.5f + rnd.Float32() * 0.5f, c = 0.5f + rnd.Float32() * 0.5f;
                    m = NewMetal(NewVec3(a, b, c), rnd.Float32() * 0.5f);
                } else {
                    m = NewDielectric(1.5f);
                }
                world->Add(NewSphere(NewVec3(cx, 0.2f, cz), 0.2f, m));
            }
        }
    }
    world->Add(NewSphere(NewVec3(0, 1, 0), 1, NewDielectric(1.5f)));
    world->Add(NewSphere(NewVec3(-4, 1, 0), 1, NewLambertian(NewSolidColor(0.4f, 0.2f, 0.1f))));
    world->Add(NewSphere(NewVec3(4, 1, 0), 1, NewMetal(NewVec3(0.7f, 0.6f, 0.5f), 0)));
    auto tree = NewBVHFromWorld(world);

    std::ofstream f(out, std::ios::binary);
    if (!f) {
        fprintf(stderr, "cannot open %s\n", out);
        return 2;
    }
    std::string err = camera->Render(tree, f, false, ends_with_png(out));
    if (!err.empty()) { // main.go:74-76 panics; a CLI reports and exits non-zero
        fprintf(stderr, "render failed: %s\n", err.c_str());
        return 1;
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const rt_stats &st = camera->last_stats;
    printf("Finished in: %.3fs  (%d spheres, %dx%d, %d spp: %.1f Msamples/s device, %.1f Mrays/s)\n", secs,
           (int)world->hittables.size(), camera->c.width, camera->c.height, spp,
           st.samples / (st.ms_render * 1e3), st.rays / (st.ms_render * 1e3));
    return 0;
}
