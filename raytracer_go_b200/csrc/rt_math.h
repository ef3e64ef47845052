// rt_math.h — float32 vector arithmetic in the reference's operation order (vec3.go).
//
// Every function is __host__ __device__ so that the per-ray logic can also be compiled by g++ for
// the pre-GPU logic checks in tests/hostsim (test-only; the product library has no CPU path).
// The translation units that include this header are compiled with -fmad=false (nvcc) /
// -ffp-contract=off (g++): `a*b + c` stays two rounded operations exactly as gc emits them on
// amd64.  Fused multiply-adds appear only where spelled fmaf() (BVH box culling,
// which only has to be conservative, never exact).
#ifndef RT_MATH_H
#define RT_MATH_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

struct V3 {
    float x, y, z;
};

RT_HD V3 v3(float x, float y, float z) {
    V3 r;
    r.x = x, r.y = y, r.z = z;
    return r;
}
RT_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); } // vec3.go:43
RT_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); } // vec3.go:67
RT_HD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); } // vec3.go:55
RT_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }     // vec3.go:91
// vec3.go:115-117 and 137-139: (x*x + y*y) + z*z, left to right, unfused
RT_HD float lensq(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
RT_HD float dot(V3 l, V3 r) { return l.x * r.x + l.y * r.y + l.z * r.z; }
// vec3.go:129-135
RT_HD V3 cross(V3 l, V3 r) { return v3(l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x); }
// float32(math.Sqrt(float64(x))) == correctly rounded sqrtf(x) (vec3.go:105, hittables.go:108)
RT_HD float sqrt32(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}
// x / y, IEEE round-to-nearest
RT_HD float div32(float x, float y) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(x, y);
#else
    return x / y;
#endif
}
// 1 / x, IEEE round-to-nearest (bit-identical to 1.0f / x; cheaper than a general divide on the GPU)
RT_HD float rcp32(float x) {
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
// vec3.go:103-107: Scale(1/len)
RT_HD V3 unit(V3 a) {
    float l = sqrt32(lensq(a));
    return a * rcp32(l);
}
// vec3.go:212-214
RT_HD V3 reflect(V3 v, V3 n) { return v - n * (2 * dot(v, n)); }
// vec3.go:216-221.  sqrt(|1 - |perp|^2|) is float32 -> f64 abs/sqrt -> float32 == sqrtf(fabsf())
RT_HD V3 refract(V3 uv, V3 n, float eta) {
    float cos_theta = dot(uv * -1.0f, n);
    V3 perp = (uv + n * cos_theta) * eta;
    float k = sqrt32(fabsf(1.0f - lensq(perp)));
    V3 par = n * (-1 * k);
    return par + perp;
}
// math.go:20-28
RT_HD float clamp01(float v) {
    if (v < 0.0f) return 0.0f;
    if (v > 1.0f) return 1.0f;
    return v;
}
// vec3.go:168-172
RT_HD bool near_zero(V3 v) {
    const float eps = 1e-8f;
    return fabsf(v.x) < eps && fabsf(v.y) < eps && fabsf(v.z) < eps;
}

RT_HD float as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union {
        uint32_t u;
        float f;
    } c;
    c.u = u;
    return c.f;
#endif
}
RT_HD uint32_t as_uint(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union {
        uint32_t u;
        float f;
    } c;
    c.f = f;
    return c.u;
#endif
}

#endif // RT_MATH_H
