// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or called by the product
// (pbrs_b200/), only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.
//
// CPU restatement (C++17, strict FP32, no FMA contraction) of the scalar / vector / colour
// arithmetic of plumer/pbrs.  Each function cites the reference file:line it follows.
// The reference is Rust and cannot be built here (no cargo/rustc; HEAD does not compile:
// scene/src/plyloader.rs:69-256 is truncated), so this is a "port" oracle.  It is pinned by
// the reference's own known-answer tests (tests/test_oracle_kat.py); the parts no reference
// test pins (slab test, BVH build/traversal, integrators) are "parity unpinned" upstream and
// rest on transcription fidelity -- see DESIGN.md.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

constexpr float kInf = std::numeric_limits<float>::infinity();
constexpr float kEps = 1.1920929e-7f;  // f32::EPSILON
constexpr float kPi = 3.14159265358979323846f;
constexpr float kFrac1Pi = 0.318309886183790671537767526745028724f;
constexpr float kFracPi2 = 1.57079632679489661923132169163975144f;

// ---- diagnostics: reference asserts that would have fired (include/pbrs_gpu.h indices) ----
struct Diag {
    uint64_t would_panic[16] = {0};
    uint64_t n_rays_extend = 0, n_rays_shadow = 0;
    uint64_t n_nodes = 0, n_tris = 0, n_spheres = 0, n_instances = 0;
    uint64_t trav_extend[4] = {0, 0, 0, 0}, trav_shadow[4] = {0, 0, 0, 0};  // split by walk kind
    void add(const Diag &o) {
        for (int i = 0; i < 4; ++i) { trav_extend[i] += o.trav_extend[i]; trav_shadow[i] += o.trav_shadow[i]; }
        for (int i = 0; i < 16; ++i) would_panic[i] += o.would_panic[i];
        n_rays_extend += o.n_rays_extend;
        n_rays_shadow += o.n_rays_shadow;
        n_nodes += o.n_nodes;
        n_tris += o.n_tris;
        n_spheres += o.n_spheres;
        n_instances += o.n_instances;
    }
};
extern thread_local Diag *g_diag;
inline void panic_flag(int kind) {
    if (g_diag) g_diag->would_panic[kind]++;
}
enum {
    P_SPHERE_INSIDE = 0, P_TBN = 1, P_HAT = 2, P_BSDF_FRAME = 3, P_MESH_UV = 4,
    P_EMPTY_BXDFS = 5, P_LOG_SAMPLE = 6, P_FRESNEL = 7, P_LAMBERT_WO = 8, P_PERLIN = 9,
    P_REFRACT = 10, P_MISC = 11, P_STACK = 12 /* GPU side only */, P_QUAD = 13
};

// ---- Rust f32 method semantics ----
inline float f_max(float a, float b) { return std::fmax(a, b); }  // f32::max (NaN-ignoring)
inline float f_min(float a, float b) { return std::fmin(a, b); }
inline float f_clamp(float x, float lo, float hi) {  // f32::clamp
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}
inline float f_signum(float x) {  // f32::signum
    if (std::isnan(x)) return x;
    return std::signbit(x) ? -1.0f : 1.0f;
}
inline float f_fract(float x) { return x - std::trunc(x); }
inline float f_recip(float x) { return 1.0f / x; }
// compiler-rt __powisf2: square-and-multiply
inline float f_powi(float a, int b) {
    const bool recip = b < 0;
    float r = 1.0f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}
// math/src/float.rs:116-122
inline float weak_recip(float x) { return x == 0.0f ? 0.0f : 1.0f / x; }
// math/src/float.rs:78-80
inline float cathetus(float h, float o) {
    return std::sqrt(f_max(f_powi(h, 2) - f_powi(o, 2), 0.0f));
}
// glam Vec3A (SSE2) lane semantics: _mm_min_ps(a,b) = a < b ? a : b (returns b on NaN)
inline float sse_min(float a, float b) { return a < b ? a : b; }
inline float sse_max(float a, float b) { return a > b ? a : b; }

// ---- Vec3 / Point3: math/src/hcm.rs:23-34 (one struct; the arithmetic is identical) ----
struct V3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float &at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // hcm.rs:170-175
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // :193-198
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }                      // :199-204
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }    // :227-232
inline V3 operator*(float s, V3 a) { return a * s; }                          // :233-238
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }    // :239-244
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }    // :86-88
inline V3 cross(V3 a, V3 v) {                                                 // :89-98
    return {a.y * v.z - a.z * v.y, a.z * v.x - a.x * v.z, a.x * v.y - a.y * v.x};
}
inline float norm_squared(V3 a) { return dot(a, a); }
inline float norm(V3 a) { return std::sqrt(norm_squared(a)); }
inline bool has_nan(V3 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }
// hcm.rs:112-117 (asserts norm2 != 0 && finite)
inline V3 hat(V3 a) {
    float n2 = norm_squared(a);
    if (!(n2 != 0.0f && std::isfinite(n2))) panic_flag(P_HAT);
    float inv = 1.0f / norm(a);
    return a * inv;
}
// hcm.rs:118-121
inline bool try_hat(V3 a, V3 *out) {
    float inv = 1.0f / norm(a);
    if (std::isfinite(inv) && inv != 0.0f) {
        *out = inv * a;
        return true;
    }
    return false;
}
// hcm.rs:124-130
inline V3 facing(V3 self, V3 normal) { return std::signbit(dot(self, normal)) ? self : -self; }
// hcm.rs:144-146
inline V3 projected_onto(V3 self, V3 other) {
    return dot(self, other) * other / norm_squared(other);
}
// hcm.rs:149-154
inline int abs_min_dimension(V3 a) {
    float ab[3] = {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)};
    int res = ab[0] < ab[1] ? 0 : 1;
    res = ab[res] < ab[2] ? res : 2;
    return res;
}
// hcm.rs:156-163
inline int max_dimension(V3 a) {
    int res = a.x > a.y ? 0 : 1;
    return a[2] > a[res] ? 2 : res;
}
inline float distance_to(V3 a, V3 b) { return norm(a - b); }
inline float squared_distance_to(V3 a, V3 b) { return norm_squared(a - b); }

// Mat3 (columns): hcm.rs:357-453
struct M3 {
    V3 c[3];
};
inline V3 operator*(const M3 &m, V3 v) { return m.c[0] * v[0] + m.c[1] * v[1] + m.c[2] * v[2]; }

// hcm.rs:595-605
inline void make_coord_system(V3 v, V3 *o1, V3 *o2) {
    int i0 = abs_min_dimension(v);
    int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
    V3 v1{0, 0, 0};
    v1.at(i1) = v[i2];
    v1.at(i2) = -v[i1];
    if (!(std::fabs(dot(v1, v)) < kEps)) panic_flag(P_MISC);
    V3 v2 = cross(v, v1);
    *o1 = hat(v1);
    *o2 = hat(v2);
}
// hcm.rs:607-611
inline V3 reflect(V3 normal, V3 wi) {
    V3 perp = dot(wi, normal) * normal / norm_squared(normal);
    V3 parallel = wi - perp;
    return wi - 2.0f * parallel;
}
// hcm.rs:625-640; returns true = Transmit, false = FullReflect
inline bool refract(V3 normal, V3 wi, float ni_over_no, V3 *out) {
    wi = hat(wi);
    normal = hat(normal);
    float cos_theta_i = dot(wi, normal);
    if (cos_theta_i < 0.0f) panic_flag(P_REFRACT);
    float sin2_theta_i = f_max(1.0f - f_powi(cos_theta_i, 2), 0.0f);
    float sin2_theta_o = sin2_theta_i * f_powi(ni_over_no, 2);
    if (sin2_theta_o >= 1.0f) {
        *out = reflect(normal, wi);
        return false;
    }
    float cos_theta_o = std::sqrt(1.0f - sin2_theta_o);
    *out = ni_over_no * -wi + (ni_over_no * cos_theta_i - cos_theta_o) * normal;
    return true;
}
// hcm.rs:647-650 (Q3: names swapped -- x uses sin(phi), y uses cos(phi))
inline V3 spherical_direction(float sin_theta, float cos_theta, float phi) {
    float cos_phi = std::sin(phi), sin_phi = std::cos(phi);
    return {sin_theta * cos_phi, sin_theta * sin_phi, cos_theta};
}
// math/src/float.rs:37-50
template <class T>
inline T barycentric_lerp(T a, T b, T c, float bc0, float bc1) {
    return (a - c) * bc0 + (b - c) * bc1 + c;
}

// ---- Color: radiometry/src/color.rs ----
struct Color {
    float r, g, b;
};
inline Color rgb(float r, float g, float b) { return {r, g, b}; }
inline Color gray(float l) { return {l, l, l}; }
inline Color black() { return {0, 0, 0}; }
inline Color operator+(Color a, Color b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }  // :121
inline Color operator-(Color a, Color b) { return {a.r - b.r, a.g - b.g, a.b - b.b}; }  // :136
inline Color operator*(Color a, float s) { return {a.r * s, a.g * s, a.b * s}; }        // :143
inline Color operator*(float s, Color a) { return a * s; }                              // :150
inline Color operator*(Color a, Color b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }  // :157
inline bool is_black(Color c) { return c.r <= 0.0f && c.g <= 0.0f && c.b <= 0.0f; }     // :57-59
inline bool is_finite(Color c) {
    return std::isfinite(c.r) && std::isfinite(c.g) && std::isfinite(c.b);
}
inline Color cw_div(Color a, Color b) { return {a.r / b.r, a.g / b.g, a.b / b.b}; }  // :97-99
inline Color cw_sqrt(Color a) { return {std::sqrt(a.r), std::sqrt(a.g), std::sqrt(a.b)}; }
inline Color cw_max(Color a, float x) { return {f_max(a.r, x), f_max(a.g, x), f_max(a.b, x)}; }
// :116-118, :222-228
inline float luminance(Color c) { return 0.21267127f * c.r + 0.71515972f * c.g + 0.07216883f * c.b; }
// :49-51  Color::rgb(u8,u8,u8)
inline Color rgb8(uint8_t r, uint8_t g, uint8_t b) {
    return {(float)r / 255.0f, (float)g / 255.0f, (float)b / 255.0f};
}

// ---- Prob: math/src/prob.rs ----
struct Prob {
    bool is_mass;
    float v;
};
inline Prob Mass(float m) { return {true, m}; }
inline Prob Density(float d) { return {false, d}; }
inline float density(Prob p) { return p.is_mass ? 0.0f : p.v; }
inline float mass(Prob p) { return p.is_mass ? p.v : 0.0f; }
inline bool is_positive(Prob p) { return p.v > 0.0f; }
inline bool is_zero(Prob p) { return p.v == 0.0f; }

// ---- sampler (DESIGN.md "Sampler"): counter-based, one u32 per
//      (seed, pixel, sample, dimension); documented deviation from the reference's OS-seeded
//      rand::thread_rng streams, preserving draw order and count ----
inline uint32_t sampler_u32(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)pixel + 1ull);
    z ^= ((uint64_t)sample + 1ull) * 0xD1B54A32D192ED03ull;
    z += ((uint64_t)dim + 1ull) * 0x8CB92BA72F3D8DD7ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
// rand 0.8 Standard f32: (u32 >> 8) * 2^-24
inline float u32_to_f32(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

}  // namespace orc
