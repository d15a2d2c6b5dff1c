// Device-side scalar / vector / colour arithmetic for the sm_100a kernels.
//
// Parity contract: every function here performs the same IEEE FP32 operations in the same order
// as the reference function it cites (compiled with -fmad=false, IEEE div/sqrt, no flush to
// zero), so that the integer outcomes of traversal (which box passes, which primitive wins) are
// bit-identical to the CPU path.  Transcendentals (sinf, logf, ...) are CUDA's libdevice
// versions: they may differ from glibc in the last ulp and are only used in shading.
//
// The headers device_*.cuh are plain C++ besides the PB_DEV qualifier, so that tests/hostsim can
// compile the very same per-path stage functions with g++ (-ffp-contract=off) and single-step
// them against the oracle without a GPU.  That build is test tooling only: the product library
// contains no host execution path for them.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

// PB_CALL marks the large shading functions that are kept as real calls on the device: the shade
// kernel is instruction-cache bound when everything is inlined (ncu: stall_no_instruction dominant,
// profiles/r1_v0_*), and one copy of each lobe / transcendental routine fixes that.
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define PB_DEV __host__ __device__ __forceinline__
#define PB_CALL __host__ __device__ __noinline__
// Functions whose struct arguments (Isect, AreaLight, out-parameters) live in local memory for the
// duration of an out-of-line call: inlined (0 = keep them out of line).  Measured: shade -8 % on C4,
// -13 % on C5, -18 % on C1 for the price of ~10 % more SASS (profiles/r2_exp_shade_inline_gridconstant.log).
#ifndef PBRS_INLINE_RECON
#define PBRS_INLINE_RECON 1
#endif
#ifndef PBRS_INLINE_AREA
#define PBRS_INLINE_AREA 1
#endif
#if PBRS_INLINE_RECON
#define PB_CALL_RECON PB_DEV
#else
#define PB_CALL_RECON PB_CALL
#endif
#if PBRS_INLINE_AREA
#define PB_CALL_AREA PB_DEV
#else
#define PB_CALL_AREA PB_CALL
#endif
#else
#define PB_DEV inline
#define PB_CALL inline
#define PB_CALL_RECON inline
#define PB_CALL_AREA inline
#endif

namespace pbrs {

constexpr float kEps = 1.1920929e-7f;  // f32::EPSILON
constexpr float kPi = 3.14159265358979323846f;
constexpr float kFrac1Pi = 0.318309886183790671537767526745028724f;  // std::f32::consts::FRAC_1_PI
constexpr float kInvPi = 0.318309886183790671537767526745028724f;
constexpr float kHalfPi = 1.57079632679489661923132169163975144f;
#define PB_INF (__builtin_huge_valf())

struct vec3 {
    float x, y, z;
};
PB_DEV vec3 mk(float x, float y, float z) { vec3 v; v.x = x; v.y = y; v.z = z; return v; }
PB_DEV float comp(vec3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
// math/src/hcm.rs:170-244
PB_DEV vec3 operator+(vec3 a, vec3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
PB_DEV vec3 operator-(vec3 a, vec3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
PB_DEV vec3 operator-(vec3 a) { return mk(-a.x, -a.y, -a.z); }
PB_DEV vec3 operator*(vec3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
PB_DEV vec3 operator*(float s, vec3 a) { return mk(a.x * s, a.y * s, a.z * s); }
PB_DEV vec3 operator/(vec3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
PB_DEV float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // hcm.rs:86-88
PB_DEV vec3 cross(vec3 a, vec3 v) {                                             // hcm.rs:89-98
    return mk(a.y * v.z - a.z * v.y, a.z * v.x - a.x * v.z, a.x * v.y - a.y * v.x);
}
PB_DEV float len2(vec3 a) { return dot(a, a); }
PB_DEV float len(vec3 a) { return sqrtf(len2(a)); }
PB_DEV bool is_nan(float x) { return x != x; }
PB_DEV bool is_fin(float x) { return fabsf(x) < PB_INF; }   // false for NaN and +-inf
PB_DEV bool is_inf(float x) { return fabsf(x) == PB_INF; }
PB_DEV bool any_nan(vec3 a) { return is_nan(a.x) || is_nan(a.y) || is_nan(a.z); }
PB_DEV uint32_t f2u(float x) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
PB_DEV float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float x; memcpy(&x, &u, 4); return x;
#endif
}
PB_DEV bool sign_neg(float x) { return (f2u(x) >> 31) != 0u; }

// Transcendentals.  The reference calls glibc's float functions (Rust f32::sin etc. lower to
// libm), which are correctly rounded for all but a tiny fraction of inputs.  CUDA's float
// versions differ from them in the last ulp far more often, and an ulp can flip a discrete
// decision downstream (a Fresnel coin, a shadow test), so shading evaluates them in FP64 and
// rounds once: this agrees with glibc in all but ~1e-3 of calls.  B200 has full-rate-class
// FP64 and these sit only in shading, never in traversal.
PB_CALL float t_sin(float x) { return (float)sin((double)x); }
PB_CALL float t_cos(float x) { return (float)cos((double)x); }
PB_CALL float t_tan(float x) { return (float)tan((double)x); }
PB_CALL float t_atan(float x) { return (float)atan((double)x); }
PB_CALL float t_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
PB_CALL float t_acos(float x) { return (float)acos((double)x); }
PB_CALL float t_log(float x) { return (float)log((double)x); }
PB_CALL float t_exp(float x) { return (float)exp((double)x); }
PB_DEV float t_hypot(float x, float y) { return (float)sqrt((double)x * (double)x + (double)y * (double)y); }

// per-thread diagnostics: reference asserts that would have fired (bit k = kind k)
struct Diag {
    uint32_t panics;
};
// (Keeping the bits in a shared-memory word per thread instead of the object -- which out-of-line
// callees pin to local memory -- was measured: no gain in the shade kernels, +1.5 % in the traversal
// kernels; profiles/r2_exp_shade_inline_gridconstant.log)
PB_DEV void flag(Diag &d, int kind) { d.panics |= 1u << kind; }
enum {
    P_SPHERE_INSIDE = 0, P_TBN = 1, P_HAT = 2, P_BSDF_FRAME = 3, P_MESH_UV = 4, P_EMPTY_BXDFS = 5,
    P_LOG_SAMPLE = 6, P_FRESNEL = 7, P_LAMBERT_WO = 8, P_PERLIN = 9, P_REFRACT = 10, P_MISC = 11,
    P_STACK = 12,  // a traversal stack overflowed (never with scenes pbrs_scene_commit accepts)
    P_QUAD = 13    // ParallelQuad's accurate-vs-coarse hit assert, shape/src/simple.rs:140-147
};

// Vec3::hat, hcm.rs:112-117
PB_DEV vec3 hat(vec3 a, Diag &d) {
    float n2 = len2(a);
    if (!(n2 != 0.0f && is_fin(n2))) flag(d, P_HAT);
    float inv = 1.0f / len(a);
    return a * inv;
}
// Vec3::try_hat, hcm.rs:118-121
PB_DEV bool try_hat(vec3 a, vec3 &out) {
    float inv = 1.0f / len(a);
    if (is_fin(inv) && inv != 0.0f) { out = inv * a; return true; }
    return false;
}
// Vec3::facing, hcm.rs:124-130
PB_DEV vec3 facing(vec3 self, vec3 normal) { return sign_neg(dot(self, normal)) ? self : -self; }
// hcm.rs:144-146
PB_DEV vec3 projected_onto(vec3 self, vec3 other) { return dot(self, other) * other / len2(other); }
// hcm.rs:149-154
PB_DEV int abs_min_dimension(vec3 a) {
    float ax = fabsf(a.x), ay = fabsf(a.y), az = fabsf(a.z);
    int res = ax < ay ? 0 : 1;
    float rv = res == 0 ? ax : ay;
    return rv < az ? res : 2;
}
PB_DEV void set_comp(vec3 &v, int i, float val) { if (i == 0) v.x = val; else if (i == 1) v.y = val; else v.z = val; }
// hcm.rs:595-605
PB_DEV void make_coord_system(vec3 v, vec3 &o1, vec3 &o2, Diag &d) {
    int i0 = abs_min_dimension(v);
    int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
    vec3 v1 = mk(0.0f, 0.0f, 0.0f);
    set_comp(v1, i1, comp(v, i2));
    set_comp(v1, i2, -comp(v, i1));
    if (!(fabsf(dot(v1, v)) < kEps)) flag(d, P_MISC);
    vec3 v2 = cross(v, v1);
    o1 = hat(v1, d);
    o2 = hat(v2, d);
}
// hcm.rs:607-611
PB_DEV vec3 reflect(vec3 normal, vec3 wi) {
    vec3 perp = dot(wi, normal) * normal / len2(normal);
    vec3 parallel = wi - perp;
    return wi - 2.0f * parallel;
}
// f32::powi = compiler-rt __powisf2 (square and multiply), SURVEY Q5
PB_DEV float sq(float x) { return x * x; }
PB_DEV float powi(float a, int b) {
    float r = 1.0f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}
// hcm.rs:625-640; true = Transmit
PB_DEV bool refract(vec3 normal, vec3 wi, float ni_over_no, vec3 &out, Diag &d) {
    wi = hat(wi, d);
    normal = hat(normal, d);
    float cos_i = dot(wi, normal);
    if (cos_i < 0.0f) flag(d, P_REFRACT);
    float sin2_i = fmaxf(1.0f - sq(cos_i), 0.0f);
    float sin2_o = sin2_i * sq(ni_over_no);
    if (sin2_o >= 1.0f) { out = reflect(normal, wi); return false; }
    float cos_o = sqrtf(1.0f - sin2_o);
    out = ni_over_no * -wi + (ni_over_no * cos_i - cos_o) * normal;
    return true;
}
// hcm.rs:647-650 (Q3: x takes sin(phi), y takes cos(phi))
PB_DEV vec3 spherical_direction(float sin_theta, float cos_theta, float phi) {
    float s = t_sin(phi), c = t_cos(phi);
    return mk(sin_theta * s, sin_theta * c, cos_theta);
}
// math/src/float.rs:37-50
PB_DEV vec3 bary_lerp(vec3 a, vec3 b, vec3 c, float bc0, float bc1) { return (a - c) * bc0 + (b - c) * bc1 + c; }
PB_DEV float bary_lerp(float a, float b, float c, float bc0, float bc1) { return (a - c) * bc0 + (b - c) * bc1 + c; }
// math/src/float.rs:116-122
PB_DEV float weak_recip(float x) { return x == 0.0f ? 0.0f : 1.0f / x; }
PB_DEV float clampf(float x, float lo, float hi) { if (x < lo) x = lo; if (x > hi) x = hi; return x; }  // f32::clamp
PB_DEV float signumf(float x) { return is_nan(x) ? x : (sign_neg(x) ? -1.0f : 1.0f); }
PB_DEV float fractf(float x) { return x - truncf(x); }
// glam Vec3A (SSE2) lane min/max: second operand on NaN
PB_DEV float lane_min(float a, float b) { return a < b ? a : b; }
PB_DEV float lane_max(float a, float b) { return a > b ? a : b; }

// ---- Color: radiometry/src/color.rs ----
struct color {
    float r, g, b;
};
PB_DEV color mkc(float r, float g, float b) { color c; c.r = r; c.g = g; c.b = b; return c; }
PB_DEV color grayc(float l) { return mkc(l, l, l); }
PB_DEV color blackc() { return mkc(0.0f, 0.0f, 0.0f); }
PB_DEV color operator+(color a, color b) { return mkc(a.r + b.r, a.g + b.g, a.b + b.b); }
PB_DEV color operator-(color a, color b) { return mkc(a.r - b.r, a.g - b.g, a.b - b.b); }
PB_DEV color operator*(color a, float s) { return mkc(a.r * s, a.g * s, a.b * s); }
PB_DEV color operator*(float s, color a) { return mkc(a.r * s, a.g * s, a.b * s); }
PB_DEV color operator*(color a, color b) { return mkc(a.r * b.r, a.g * b.g, a.b * b.b); }
PB_DEV bool is_black(color c) { return c.r <= 0.0f && c.g <= 0.0f && c.b <= 0.0f; }  // color.rs:57-59
PB_DEV bool is_finite(color c) { return is_fin(c.r) && is_fin(c.g) && is_fin(c.b); }
PB_DEV color cw_div(color a, color b) { return mkc(a.r / b.r, a.g / b.g, a.b / b.b); }
PB_DEV color cw_sqrt(color a) { return mkc(sqrtf(a.r), sqrtf(a.g), sqrtf(a.b)); }
PB_DEV color cw_max(color a, float x) { return mkc(fmaxf(a.r, x), fmaxf(a.g, x), fmaxf(a.b, x)); }
PB_DEV float luminance(color c) { return 0.21267127f * c.r + 0.71515972f * c.g + 0.07216883f * c.b; }  // :222-228

// ---- sampler (DESIGN.md "Sampler"): one u32 per (seed, pixel, sample, dimension) ----
PB_DEV uint32_t sampler_u32(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)pixel + 1ull);
    z ^= ((uint64_t)sample + 1ull) * 0xD1B54A32D192ED03ull;
    z += ((uint64_t)dim + 1ull) * 0x8CB92BA72F3D8DD7ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
// rand 0.8 `Standard` for f32: (u32 >> 8) * 2^-24
PB_DEV float u32_to_unit(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

struct Sampler {
    uint64_t seed;
    uint32_t pixel, sample;
    PB_DEV float f(uint32_t dim) const { return u32_to_unit(sampler_u32(seed, pixel, sample, dim)); }
    PB_DEV uint32_t u(uint32_t dim) const { return sampler_u32(seed, pixel, sample, dim); }
};

// ---- Ray / Interaction ----
struct Ray {
    vec3 o, d;
    float t_max;
};
// geometry/src/ray.rs:40-46
PB_DEV bool in_extent(float t, float t_max) { return !(t < kEps || t >= t_max); }
PB_DEV vec3 at(const Ray &r, float t) { return r.o + t * r.d; }

// geometry/src/interaction.rs:12-21
struct Isect {
    vec3 pos;
    float t;
    float u, v;
    vec3 normal, wo;
    vec3 tangent;  // tbn_frame.cols[0]
};
// interaction.rs:23-33
PB_DEV Isect isect_new(vec3 pos, float t, float u, float v, vec3 normal, vec3 wo, Diag &d) {
    if (!(dot(normal, wo) >= 0.0f)) flag(d, P_SPHERE_INSIDE);
    Isect i;
    i.pos = pos; i.t = t; i.u = u; i.v = v; i.normal = normal; i.wo = wo;
    i.tangent = mk(0.0f, 0.0f, 0.0f);
    return i;
}
// interaction.rs:45-61 (only the tangent column is consumed downstream: transform.rs:316, bsdf.rs:20)
PB_DEV void with_dpdu(Isect &i, vec3 dpdu, Diag &d) {
    if (!(fabsf(dot(i.normal, dpdu)) < 1e-3f)) flag(d, P_TBN);
    vec3 n = hat(i.normal, d);
    vec3 bt = hat(cross(n, dpdu), d);
    vec3 t = cross(bt, n);
    float det = dot(cross(t, bt), n);
    if (!(fabsf(det - 1.0f) < 1e-4f)) flag(d, P_TBN);
    i.tangent = t;
}
// interaction.rs:63-70
PB_DEV Ray spawn_ray(const Isect &i, vec3 dir) {
    vec3 out_n = signumf(dot(dir, i.normal)) * i.normal;
    Ray r;
    r.o = i.pos + out_n * 0.001f;
    r.d = dir;
    r.t_max = PB_INF;
    return r;
}
PB_DEV Ray spawn_limited_ray_to(const Isect &i, vec3 p) {
    Ray r = spawn_ray(i, p - i.pos);
    r.t_max = 1.0f - 0.001f;
    return r;
}

}  // namespace pbrs
