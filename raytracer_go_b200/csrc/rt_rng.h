// rt_rng.h — counter-based Philox4x32-7 stream, one per (pixel, sample) path.
//
// The reference draws from clock-seeded math/rand generators (camera.go:170-171, materials.go:103);
// that stream is not reproducible and not part of any contract.  Here the stream of a path is the
// sequence of BLOCKS
//   block b = Philox4x32-7( counter = (pixel, sample, b, 0), key = (seed_lo, seed_hi) ),  b = 0,1,2,...
// each giving four uniforms  Float32() = (word >> 8) * 2^-24  in [0,1), and every consumer takes
// whole blocks (a partly used block is dropped):
//   Camera.GetRay          one block = (dx, dy, disk.x, disk.y); while the disk pair is rejected
//                          (vec3.go:203-210) another block = two more candidate pairs, tried in order
//   unit-sphere rejection  one block per trial = (x, y, z, unused)      (vec3.go:182-190)
//   Dielectric.Scatter     one block, first word = the uniform of materials.go:103
// Whole-block consumption keeps the Philox rounds at warp-convergent program points (every lane
// still in a rejection loop generates together) instead of inside a per-lane "buffer empty" branch.
// Any (pixel, sample) can be generated on any GPU in any order, so sample-split / tile-split renders
// draw exactly the samples of the single-GPU render.
#ifndef RT_RNG_H
#define RT_RNG_H

#include "rt_math.h"

struct RngBlock {
    float u0, u1, u2, u3;    // Float32() of the four words
    uint32_t w0, w1, w2, w3; // the words themselves (for the fused mappings below)
};

// Rounds: Random123 (Salmon et al., SC'11, table 2) lists Philox4x32 with 7 rounds as the fastest variant that passes
// TestU01's BigCrush ("Crush-resistant"); 10 rounds, its default, add a safety margin that a Monte-Carlo integrator
// of a 405 M-sample frame does not need.  The rejection samplers make the generator 18 % of the frame's instructions
// (profiles/r02p_*_by_line.txt), 40 instructions per block at 10 rounds; 7 rounds measured +4.1 % on C2, +3.6 % on
// C3, +6.0 % on the Cornell box (profiles/r02q).  Both round counts are pinned by Random123's known-answer vectors
// (tests/test_oracle_kat.py); the oracle draws the same 7-round stream, and the converged-image test compares the
// device with an oracle frame drawn from the 10-round stream.  -DRT_PHILOX_ROUNDS=10 builds the other variant.
//
// The round keys (k0 + r*W0, k1 + r*W1) depend on the seed only.  On the device they live in
// constant memory (philox_round_keys() fills the array, the host uploads it before a launch), so
// the xor of a round takes its key as a constant-bank operand: no key registers, no 18 key adds
// per block.  The host build (tests/hostsim) keeps them in a thread-local array.
#ifndef RT_PHILOX_ROUNDS
#define RT_PHILOX_ROUNDS 7
#endif
RT_HD void philox_round_keys(uint64_t seed, uint32_t *rk) {
    uint32_t q0 = (uint32_t)seed, q1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < RT_PHILOX_ROUNDS; r++) {
        rk[2 * r] = q0, rk[2 * r + 1] = q1;
        q0 += 0x9E3779B9u;
        q1 += 0xBB67AE85u;
    }
}
#if defined(__CUDACC__)
static __constant__ uint32_t c_philox_rk[2 * RT_PHILOX_ROUNDS];
#endif
#if !defined(__CUDA_ARCH__)
static thread_local uint32_t h_philox_rk[2 * RT_PHILOX_ROUNDS];
#endif

struct PathRng {
    uint32_t pixel, sample, block;

    // `seed` must be the seed whose round keys are loaded (device: the launch's constant array)
    RT_HD void init(uint64_t seed, uint32_t pixel_, uint32_t sample_) {
        pixel = pixel_, sample = sample_, block = 0;
#if !defined(__CUDA_ARCH__)
        philox_round_keys(seed, h_philox_rk);
#else
        (void)seed;
#endif
    }
    RT_HD static float to_f32(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
    // next block of the stream as four uniforms on [0,1) (rand.Float32(), camera.go:290-291)
    RT_HD RngBlock next() {
        uint32_t c0 = pixel, c1 = sample, c2 = block, c3 = 0;
#if defined(__CUDA_ARCH__)
        const uint32_t *rk = c_philox_rk;
#else
        const uint32_t *rk = h_philox_rk;
#endif
#pragma unroll
        for (int r = 0; r < RT_PHILOX_ROUNDS; r++) {
            // one 32x32->64 multiply per half (IMAD.WIDE.U32), high word xor-ed, low word passed on
            const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
            const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk[2 * r];
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
            c0 = n0, c1 = (uint32_t)p1, c2 = n2, c3 = (uint32_t)p0;
        }
        block++;
        RngBlock b;
        b.u0 = to_f32(c0), b.u1 = to_f32(c1), b.u2 = to_f32(c2), b.u3 = to_f32(c3);
        b.w0 = c0, b.w1 = c1, b.w2 = c2, b.w3 = c3;
        return b;
    }
};

// math.go:30-32: min + r*(max-min)
RT_HD float rand_range(float r, float lo, float hi) { return lo + r * (hi - lo); }
// rand_range(Float32(), -1, 1) and -0.5 + Float32() straight from a Philox word.  Float32() is k * 2^-24 with
// k < 2^24 (exact), and so are r * 2 and k * 2^-23: the reference's "multiply, then add" rounds once, at the add,
// and the fused multiply-add below rounds the same real number once — identical bits, one instruction instead
// of three (the oracle keeps the unfused form; the parity tests compare the two).
RT_HD float rand_pm1(uint32_t w) { return fmaf((float)(w >> 8), 1.0f / 8388608.0f, -1.0f); }
RT_HD float rand_centered(uint32_t w) { return fmaf((float)(w >> 8), 1.0f / 16777216.0f, -0.5f); }

// vec3.go:182-190 (NewVec3UnitRandOnUnitSphere32): cube rejection, then Unit(); one block per trial
RT_HD V3 rand_unit(PathRng &rng) {
    V3 v;
    for (;;) {
        const RngBlock b = rng.next();
        v = v3(rand_pm1(b.w0), rand_pm1(b.w1), rand_pm1(b.w2));
        if (lensq(v) < 1.0f) break;
    }
    return unit(v); // after the loop: the sqrt and divide run once, with the lanes reconverged
}

#endif // RT_RNG_H
