// rt_rng.h — counter-based Philox4x32-10 stream, one per (pixel, sample) path.
//
// The reference draws from clock-seeded math/rand generators (camera.go:170-171, materials.go:103);
// that stream is not reproducible and not part of any contract.  Here the stream of a path is
//   words of Philox4x32-10( counter = (pixel, sample, block, 0), key = (seed_lo, seed_hi) ),
// block = 0,1,2,..., consumed in order; Float32() = (word >> 8) * 2^-24  in [0,1).
// Any (pixel, sample) can therefore be generated on any GPU in any order (sample-split /
// tile-split renders draw exactly the samples of the single-GPU render).
#ifndef RT_RNG_H
#define RT_RNG_H

#include "rt_math.h"

RT_HD uint32_t rt_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

struct PathRng {
    uint32_t pixel, sample, block;
    uint32_t k0, k1;
    uint32_t b0, b1, b2, b3; // unread words of the current block, b0 next
    uint32_t avail;

    RT_HD void init(uint64_t seed, uint32_t pixel_, uint32_t sample_) {
        pixel = pixel_, sample = sample_, block = 0;
        k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        avail = 0;
        b0 = b1 = b2 = b3 = 0;
    }
    RT_HD void refill() {
        uint32_t c0 = pixel, c1 = sample, c2 = block, c3 = 0;
        uint32_t q0 = k0, q1 = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0 = rt_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = rt_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ q0;
            uint32_t n2 = hi0 ^ c3 ^ q1;
            c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
            q0 += 0x9E3779B9u;
            q1 += 0xBB67AE85u;
        }
        b0 = c0, b1 = c1, b2 = c2, b3 = c3;
        block++;
        avail = 4;
    }
    RT_HD uint32_t u32() {
        if (avail == 0) refill();
        uint32_t r = b0;
        b0 = b1, b1 = b2, b2 = b3;
        avail--;
        return r;
    }
    // uniform on [0,1) like rand.Float32() (camera.go:290-291)
    RT_HD float f32() { return (float)(u32() >> 8) * (1.0f / 16777216.0f); }
    // math.go:30-32
    RT_HD float range(float lo, float hi) { return lo + f32() * (hi - lo); }
};

// vec3.go:182-190 (NewVec3UnitRandOnUnitSphere32): cube rejection, then Unit()
RT_HD V3 rand_unit(PathRng &rng) {
    for (;;) {
        float x = rng.range(-1.0f, 1.0f);
        float y = rng.range(-1.0f, 1.0f);
        float z = rng.range(-1.0f, 1.0f);
        V3 v = v3(x, y, z);
        if (lensq(v) < 1.0f) return unit(v);
    }
}

#endif // RT_RNG_H
