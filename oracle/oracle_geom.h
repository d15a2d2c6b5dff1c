// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_math.h header).
// Restates the `geometry` crate: ray, bbox, interaction, camera, transform, bxdf, microfacet.
#pragma once
#include "oracle_math.h"

namespace orc {

// ---- Ray: geometry/src/ray.rs:17-51 ----
struct Ray {
    V3 origin, dir;
    float t_max;
};
inline Ray make_ray(V3 o, V3 d) { return Ray{o, d, kInf}; }
// ray.rs:40-46; returns false for None
inline bool truncated_t(const Ray &r, float t) { return !(t < kEps || t >= r.t_max); }
inline V3 position_at(const Ray &r, float t) { return r.origin + t * r.dir; }

// ---- BBox: geometry/src/bvh.rs:11-143 (glam::Vec3A lanes => SSE min/max semantics) ----
struct BBox {
    float mn[3], mx[3];
};
inline BBox bbox_empty() { return BBox{{kInf, kInf, kInf}, {-kInf, -kInf, -kInf}}; }
// bvh.rs:28-35
inline BBox bbox_new(V3 p0, V3 p1) {
    BBox b;
    for (int i = 0; i < 3; ++i) {
        b.mn[i] = sse_min(p0[i], p1[i]);
        b.mx[i] = sse_max(p0[i], p1[i]);
    }
    return b;
}
// bvh.rs:37-44 (scalar f32::min / f32::max)
inline BBox bbox_union_pt(BBox b, V3 p) {
    BBox r = b;
    for (int i = 0; i < 3; ++i) {
        r.mn[i] = f_min(b.mn[i], p[i]);
        r.mx[i] = f_max(b.mx[i], p[i]);
    }
    return r;
}
// bvh.rs:138-143
inline BBox bbox_union(BBox a, BBox b) {
    BBox r;
    for (int i = 0; i < 3; ++i) {
        r.mn[i] = sse_min(a.mn[i], b.mn[i]);
        r.mx[i] = sse_max(a.mx[i], b.mx[i]);
    }
    return r;
}
// bvh.rs:46-49
inline V3 bbox_midpoint(const BBox &b) {
    return {(b.mx[0] - b.mn[0]) * 0.5f + b.mn[0], (b.mx[1] - b.mn[1]) * 0.5f + b.mn[1],
            (b.mx[2] - b.mn[2]) * 0.5f + b.mn[2]};
}
inline V3 bbox_diag(const BBox &b) { return {b.mx[0] - b.mn[0], b.mx[1] - b.mn[1], b.mx[2] - b.mn[2]}; }
inline V3 bbox_min(const BBox &b) { return {b.mn[0], b.mn[1], b.mn[2]}; }
// bvh.rs:75-82
inline float bbox_area(const BBox &b) {
    V3 d = bbox_diag(b);
    if (!std::signbit(d.x) && !std::signbit(d.y) && !std::signbit(d.z))
        return (d.x * d.y + d.y * d.z + d.z * d.x) * 2.0f;
    return 0.0f;
}
// bvh.rs:84-99.  Slab test with true divisions.  glam 0.20 (SSE2) semantics, restated from its
// published source: Vec3A::min/max = _mm_min_ps/_mm_max_ps (return the 2nd operand on NaN);
// max_element = max_ps(max_ps(x,z), max_ps(y,z)), min_element likewise; then f32::max/min.
inline bool bbox_intersect(const BBox &b, const Ray &r, float *t_low_out = nullptr) {
    float lo[3], hi[3];
    for (int i = 0; i < 3; ++i) {
        float t0 = (b.mn[i] - r.origin[i]) / r.dir[i];
        float t1 = (b.mx[i] - r.origin[i]) / r.dir[i];
        lo[i] = sse_min(t0, t1);
        hi[i] = sse_max(t0, t1);
    }
    float max_el = sse_max(sse_max(lo[0], lo[2]), sse_max(lo[1], lo[2]));
    float min_el = sse_min(sse_min(hi[0], hi[2]), sse_min(hi[1], hi[2]));
    float t_low = f_max(max_el, 0.0f);
    float t_high = f_min(min_el, r.t_max);
    if (t_low_out) *t_low_out = t_low;
    return t_low <= t_high;
}

// ---- Interaction: geometry/src/interaction.rs:12-70 ----
struct Interaction {
    V3 pos;
    float ray_t;
    float u, v;
    V3 normal, wo;
    M3 tbn;  // cols: tangent(dpdu), bitangent, normal-hat
};
// interaction.rs:23-33
inline Interaction isect_new(V3 pos, float t, float u, float v, V3 normal, V3 wo) {
    if (!(dot(normal, wo) >= 0.0f)) panic_flag(P_SPHERE_INSIDE);
    Interaction i;
    i.pos = pos; i.ray_t = t; i.u = u; i.v = v; i.normal = normal; i.wo = wo;
    i.tbn = M3{{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
    return i;
}
inline Interaction isect_rayless(V3 pos, float u, float v, V3 normal) {
    return isect_new(pos, 0.0f, u, v, normal, V3{0, 0, 0});
}
// interaction.rs:45-61
inline Interaction with_dpdu(Interaction self, V3 dpdu) {
    if (!(std::fabs(dot(self.normal, dpdu)) < 1e-3f)) panic_flag(P_TBN);
    V3 normal = hat(self.normal);
    V3 bitangent = hat(cross(normal, dpdu));
    V3 t = cross(bitangent, normal);
    float det = dot(cross(t, bitangent), normal);
    if (!(std::fabs(det - 1.0f) < 1e-4f)) panic_flag(P_TBN);
    self.tbn = M3{{t, bitangent, normal}};
    return self;
}
// interaction.rs:63-66
inline Ray spawn_ray(const Interaction &i, V3 dir) {
    V3 out_normal = f_signum(dot(dir, i.normal)) * i.normal;
    return make_ray(i.pos + out_normal * 0.001f, dir);
}
// interaction.rs:68-70
inline Ray spawn_limited_ray_to(const Interaction &i, V3 p) {
    Ray r = spawn_ray(i, p - i.pos);
    r.t_max = 1.0f - 0.001f;
    return r;
}

// ---- Camera: geometry/src/camera.rs:5-77 ----
struct Camera {
    V3 center, a, b, c;
    uint32_t width, height;
    M3 orientation;
};
// camera.rs:19-35 (+ look_at :37-44)
inline Camera camera_new(uint32_t width, uint32_t height, float fov_y_rad) {
    float aspect_ratio = (float)width / (float)height;
    float half_vertical = std::tan(fov_y_rad * 0.5f);
    float half_horizontal = half_vertical * aspect_ratio;
    Camera cam;
    cam.center = V3{0, 0, 0};
    cam.a = V3{half_horizontal / (float)(width / 2), 0.0f, 0.0f};
    cam.b = V3{0.0f, -half_vertical / (float)(height / 2), 0.0f};
    cam.c = V3{-half_horizontal, half_vertical, 1.0f};
    cam.width = width;
    cam.height = height;
    cam.orientation = M3{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
    return cam;
}
inline void camera_look_at(Camera &cam, V3 from, V3 target, V3 up) {
    V3 forward = hat(target - from);
    V3 right = hat(cross(up, forward));
    V3 up2 = cross(forward, right);
    cam.orientation = M3{{right, up2, forward}};
    cam.center = from;
}
// camera.rs:65-77
inline Ray shoot_ray(const Camera &cam, uint32_t row, uint32_t col, float dx, float dy) {
    float x = (float)col + f_fract(dx);
    float y = (float)row + f_fract(dy);
    V3 c = cam.orientation * cam.c;
    V3 a = cam.orientation * cam.a;
    V3 b = cam.orientation * cam.b;
    V3 dir = c + a * x + b * y;
    return make_ray(cam.center, dir);
}

// ---- AffineTransform: geometry/src/transform.rs:16-19,267-320; Mat4 ops hcm.rs:539-576 ----
struct M4 {
    float c[4][4];  // c[col][row]
};
inline M4 m4_identity() {
    M4 m;
    std::memset(&m, 0, sizeof m);
    for (int i = 0; i < 4; ++i) m.c[i][i] = 1.0f;
    return m;
}
struct Affine {
    M4 fwd, inv;
    bool is_identity;
};
// hcm.rs:539-544: ((c0*v0 + c1*v1) + c2*v2) + c3*v3, per lane, separate mul and add
inline void m4_mul_v4(const M4 &m, const float v[4], float out[4]) {
    for (int r = 0; r < 4; ++r)
        out[r] = m.c[0][r] * v[0] + m.c[1][r] * v[1] + m.c[2][r] * v[2] + m.c[3][r] * v[3];
}
// transform.rs:267-272
inline V3 affine_vec(const M4 &m, V3 x) {
    float v[4] = {x.x, x.y, x.z, 0.0f}, o[4];
    m4_mul_v4(m, v, o);
    return {o[0], o[1], o[2]};
}
// transform.rs:273-281 (asserts w == 1; try_from divides otherwise)
inline V3 affine_pt(const M4 &m, V3 p) {
    float v[4] = {p.x, p.y, p.z, 1.0f}, o[4];
    m4_mul_v4(m, v, o);
    if (o[3] != 1.0f) panic_flag(P_MISC);
    return {o[0], o[1], o[2]};
}
// transform.rs:282-286
inline Ray affine_ray(const M4 &m, const Ray &r) {
    return Ray{affine_pt(m, r.origin), affine_vec(m, r.dir), r.t_max};
}
// transform.rs:287-308
inline BBox affine_bbox(const M4 &m, const BBox &b) {
    V3 bases[3] = {{m.c[0][0], m.c[0][1], m.c[0][2]},
                   {m.c[1][0], m.c[1][1], m.c[1][2]},
                   {m.c[2][0], m.c[2][1], m.c[2][2]}};
    BBox res = bbox_empty();
    V3 diag = bbox_diag(b);
    for (int i = 0; i < 8; ++i) {
        V3 corner = affine_pt(m, bbox_min(b));
        if (i & 1) corner = corner + diag[0] * bases[0];
        if (i & 2) corner = corner + diag[1] * bases[1];
        if (i & 4) corner = corner + diag[2] * bases[2];
        res = bbox_union_pt(res, corner);
    }
    return res;
}
// hcm.rs:558-564 with the transposed inverse: (inv^T) * v, three terms only
inline V3 m4_transpose_mul_v3(const M4 &m, V3 v) {
    // column j of m^T is row j of m: (m.c[0][j], m.c[1][j], m.c[2][j], m.c[3][j])
    V3 c0{m.c[0][0], m.c[1][0], m.c[2][0]};
    V3 c1{m.c[0][1], m.c[1][1], m.c[2][1]};
    V3 c2{m.c[0][2], m.c[1][2], m.c[2][2]};
    return c0 * v[0] + c1 * v[1] + c2 * v[2];
}
// transform.rs:309-320
inline Interaction affine_isect(const Affine &t, const Interaction &i) {
    V3 new_pos = affine_pt(t.fwd, i.pos);
    V3 new_wo = affine_vec(t.fwd, i.wo);
    V3 new_normal = m4_transpose_mul_v3(t.inv, i.normal);
    Interaction res = isect_new(new_pos, i.ray_t, i.u, i.v, new_normal, new_wo);
    return with_dpdu(res, affine_vec(t.fwd, i.tbn.c[0]));
}

// ---- Omega: geometry/src/bxdf.rs:42-177 (local shading frame, +Z = normal) ----
inline float cos_theta(V3 w) { return w.z; }
inline float cos2_theta(V3 w) { return f_powi(w.z, 2); }
inline float sin2_theta(V3 w) { return 1.0f - cos2_theta(w); }
inline float sin_theta(V3 w) { return std::sqrt(f_max(sin2_theta(w), 0.0f)); }
inline float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
inline float try_divide_or(float x, float d, float fallback) { return d == 0.0f ? fallback : x / d; }
inline float cos_phi(V3 w) { return try_divide_or(w.x, std::hypot(w.x, w.y), 1.0f); }
inline float sin_phi(V3 w) { return try_divide_or(w.y, std::hypot(w.x, w.y), 0.0f); }
inline float cos2_phi(V3 w) { return try_divide_or(w.x * w.x, w.x * w.x + w.y * w.y, 1.0f); }
inline float sin2_phi(V3 w) { return try_divide_or(w.y * w.y, w.x * w.x + w.y * w.y, 0.0f); }
// bxdf.rs:85-93 (returns (x/h, y/h) -- named sin,cos upstream)
inline void sin_cos_phi(V3 w, float *s, float *c) {
    float h = std::hypot(w.x, w.y);
    if (h == 0.0f) { *s = 0.0f; *c = 1.0f; }
    else { *s = w.x / h; *c = w.y / h; }
}
inline bool same_hemisphere(V3 a, V3 b) { return a.z * b.z >= 0.0f; }  // bxdf.rs:111-113
inline bool bisector(V3 a, V3 b, V3 *out) { return try_hat(a + b, out); }  // :143-146
inline V3 face_forward(V3 w, V3 n) { return std::signbit(dot(w, n)) ? -w : w; }  // :149-155

// bxdf.rs:187-200
inline void concentric_sample_disk(float u, float v, float *ox, float *oy) {
    float x = u * 2.0f - 1.0f;
    float y = v * 2.0f - 1.0f;
    if (x == 0.0f && y == 0.0f) { *ox = 0; *oy = 0; return; }
    float r = std::fabs(std::fabs(x) > std::fabs(y) ? x : y);
    float hyp = std::hypot(x, y);
    float ct = x / hyp, st = y / hyp;
    *ox = r * ct;
    *oy = r * st;
}
// bxdf.rs:202-206
inline V3 cos_sample_hemisphere(float u, float v) {
    float x, y;
    concentric_sample_disk(u, v, &x, &y);
    float z = std::sqrt(f_max(1.0f - x * x - y * y, 0.0f));
    return {x, y, z};
}
inline float cos_hemisphere_pdf(V3 w) { return w.z * kFrac1Pi; }  // bxdf.rs:208-210

// ---- Fresnel: bxdf.rs:284-393 ----
enum FresnelKind { FR_NOP = 0, FR_DIELECTRIC = 1, FR_CONDUCTOR = 2 };
struct Fresnel {
    int kind;
    float eta_front, eta_back;  // dielectric
    Color eta_i, eta_t, k;      // conductor
};
inline Fresnel fresnel_nop() { return Fresnel{FR_NOP, 0, 0, black(), black(), black()}; }
inline Fresnel fresnel_dielectric(float f, float b) {
    return Fresnel{FR_DIELECTRIC, f, b, black(), black(), black()};
}
inline Fresnel fresnel_conductor(Color eta_real, Color eta_imag) {  // bxdf.rs:299-305
    return Fresnel{FR_CONDUCTOR, 0, 0, gray(1.0f), eta_real, eta_imag};
}
// bxdf.rs:308-342
inline float fresnel_refl_coeff(const Fresnel &f, float cos_theta_i) {
    if (f.kind == FR_NOP) return 1.0f;
    if (f.kind == FR_CONDUCTOR) { panic_flag(P_FRESNEL); return 1.0f; }
    cos_theta_i = f_clamp(cos_theta_i, -1.0f, 1.0f);
    float eta_i, eta_t;
    if (cos_theta_i > 0.0f) { eta_i = f.eta_front; eta_t = f.eta_back; }
    else { eta_i = f.eta_back; eta_t = f.eta_front; cos_theta_i = -cos_theta_i; }
    float sin_theta_i = std::sqrt(f_max(1.0f - f_powi(cos_theta_i, 2), 0.0f));
    float sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1.0f) return 1.0f;
    float cos_theta_t = std::sqrt(f_max(1.0f - f_powi(sin_theta_t, 2), 0.0f));
    float r_perp = (eta_i * cos_theta_i - eta_t * cos_theta_t) / (eta_i * cos_theta_i + eta_t * cos_theta_t);
    float r_par = (eta_t * cos_theta_i - eta_i * cos_theta_t) / (eta_t * cos_theta_i + eta_i * cos_theta_t);
    return (f_powi(r_par, 2) + f_powi(r_perp, 2)) * 0.5f;
}
// bxdf.rs:344-392
inline Color fresnel_eval(const Fresnel &f, float cos_theta_i) {
    if (f.kind != FR_CONDUCTOR) return gray(fresnel_refl_coeff(f, cos_theta_i));
    Color eta = cw_div(f.eta_t, f.eta_i);
    Color eta2 = eta * eta;
    Color etak = cw_div(f.k, f.eta_i);
    Color etak2 = etak * etak;
    float cos2 = f_powi(f_clamp(cos_theta_i, -1.0f, 1.0f), 2);
    float sin2 = 1.0f - cos2;
    Color t0 = eta2 - etak2 - gray(sin2);
    Color a2_plus_b2 = cw_sqrt(t0 * t0 + 4.0f * eta2 * etak2);
    Color t1 = a2_plus_b2 + gray(cos2);
    Color a = cw_sqrt((a2_plus_b2 + t0) * 0.5f);
    Color t2 = 2.0f * a * cos_theta_i;
    Color ratio_s = cw_div(t1 - t2, t1 + t2);
    if (!is_finite(ratio_s)) panic_flag(P_FRESNEL);
    Color t3 = cos2 * a2_plus_b2 + gray(f_powi(sin2, 2));
    Color t4 = t2 * sin2;
    Color ratio_p = ratio_s * cw_div(t3 - t4, t3 + t4);
    if (!is_finite(ratio_p)) panic_flag(P_FRESNEL);
    return cw_max((ratio_s + ratio_p) * 0.5f, 0.0f);
}

// ---- MicrofacetDistrib (Beckmann only is instantiated): geometry/src/microfacet.rs ----
// microfacet.rs:16-23 (left-to-right, not Horner: Q5)
inline float roughness_to_alpha(float roughness) {
    float x = f_max(std::log(roughness), -8.0f);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x +
           0.000640711f * x * x * x * x;
}
struct Beckmann {
    float ax, ay;
};
// microfacet.rs:36-59
inline float mf_d(const Beckmann &m, V3 wh) {
    float tan2 = tan2_theta(wh);
    float cos4 = f_powi(cos2_theta(wh), 2);
    if (std::isnan(tan2) || std::isnan(cos4)) panic_flag(P_MISC);
    if (std::isinf(tan2)) return 0.0f;
    float x = cos2_phi(wh) / f_powi(m.ax, 2) + sin2_phi(wh) / f_powi(m.ay, 2);
    return std::exp(x * -tan2) / (kPi * m.ax * m.ay * cos4);
}
// microfacet.rs:64-88
inline float mf_lambda(const Beckmann &m, V3 w) {
    float abs_tan = std::fabs(std::sqrt(tan2_theta(w)));
    if (std::isinf(abs_tan)) return 0.0f;
    float alpha = std::sqrt(cos2_phi(w) * f_powi(m.ax, 2) + sin2_phi(w) * f_powi(m.ay, 2));
    float a = f_recip(alpha * abs_tan);
    if (a >= 1.6f) return 0.0f;
    return (1.0f - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
}
inline float mf_g1(const Beckmann &m, V3 w) { return f_recip(1.0f + mf_lambda(m, w)); }
// microfacet.rs:107-109
inline float mf_g(const Beckmann &m, V3 wo, V3 wi) {
    return f_recip(1.0f + mf_lambda(m, wo) + mf_lambda(m, wi));
}
// microfacet.rs:111-124 (cfg(not(sample_visible_area)) is the live branch)
inline float mf_pdf(const Beckmann &m, V3 /*wo*/, V3 wh) {
    float x = mf_d(m, wh);
    float y = std::fabs(cos_theta(wh));
    if (std::isnan(x * y)) panic_flag(P_MISC);
    return mf_d(m, wh) * std::fabs(cos_theta(wh));
}
// microfacet.rs:126-159
inline V3 mf_sample_wh(const Beckmann &m, V3 wo, float u, float v) {
    float tan2, phi;
    if (m.ax == m.ay) {
        float log_sample = std::log(1.0f - u);
        if (!std::isfinite(log_sample)) panic_flag(P_LOG_SAMPLE);
        tan2 = -f_powi(m.ax, 2) * log_sample;
        phi = v * 2.0f * kPi;
    } else {
        float log_sample = std::log(1.0f - u);
        if (!std::isfinite(log_sample)) panic_flag(P_LOG_SAMPLE);
        phi = std::atan(m.ay / m.ax * std::tan(2.0f * kPi * v + kFracPi2));
        if (v >= 0.5f) phi += kPi;
        float sp = std::sin(phi), cp = std::cos(phi);
        float alpha2 = f_powi(cp / m.ax, 2) + f_powi(sp / m.ay, 2);
        tan2 = -log_sample / alpha2;
    }
    float ct = f_recip(std::sqrt(1.0f + tan2));
    float st = ct * std::sqrt(tan2);
    V3 wh = spherical_direction(st, ct, phi);
    return face_forward(wh, wo);
}

// ---- BxDF lobes: bxdf.rs:395-639 ----
enum LobeKind { LOBE_SPECULAR = 0, LOBE_LAMBERT = 1, LOBE_MICROFACET = 2, LOBE_OREN_NAYAR = 3 };
enum Intrusion { INTR_REFLECTION = 0, INTR_TRANSMISSION = 1, INTR_HYBRID = 2 };
struct Lobe {
    int kind;
    Color albedo;
    Fresnel fresnel;
    int intrusion;
    Beckmann distrib;
    float on_a, on_b;  // Oren-Nayar coefficients (only reachable from the KAT hooks)
};
inline Lobe lobe_lambert(Color albedo) {
    return Lobe{LOBE_LAMBERT, albedo, fresnel_nop(), 0, {0, 0}, 0, 0};
}
// bxdf.rs:528-536
inline Lobe lobe_oren_nayar(Color albedo, float sigma_rad) {
    float s2 = f_powi(sigma_rad, 2);
    float a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
    float b = 0.45f * s2 / (s2 + 0.09f);
    return Lobe{LOBE_OREN_NAYAR, albedo, fresnel_nop(), 0, {0, 0}, a, b};
}
inline Lobe lobe_mirror(Color albedo) {
    return Lobe{LOBE_SPECULAR, albedo, fresnel_nop(), INTR_REFLECTION, {0, 0}, 0, 0};
}
inline Lobe lobe_dielectric(Color albedo, float eta_outer, float eta_inner) {
    return Lobe{LOBE_SPECULAR, albedo, fresnel_dielectric(eta_outer, eta_inner), INTR_HYBRID, {0, 0}, 0, 0};
}
inline Lobe lobe_transmit(Color albedo, float eta_outer, float eta_inner) {
    return Lobe{LOBE_SPECULAR, albedo, fresnel_dielectric(eta_outer, eta_inner), INTR_TRANSMISSION, {0, 0}, 0, 0};
}
inline Lobe lobe_microfacet(Color albedo, Beckmann d, Fresnel f) {
    return Lobe{LOBE_MICROFACET, albedo, f, 0, d, 0, 0};
}

// bxdf.rs:427-434
inline void specular_reflect(const Lobe &l, V3 wo, V3 *wi, Color *c) {
    *wi = V3{-wo.x, -wo.y, wo.z};
    Color fr = fresnel_eval(l.fresnel, cos_theta(*wi));
    *c = fr * l.albedo * weak_recip(std::fabs(cos_theta(*wi)));
}
// bxdf.rs:436-454
inline void specular_refract(const Lobe &l, V3 wo, float eta_front, float eta_back, V3 *wi, Color *c) {
    float eta_i, eta_t;
    V3 normal;
    if (cos_theta(wo) > 0.0f) { eta_i = eta_front; eta_t = eta_back; normal = V3{0, 0, 1}; }
    else { eta_i = eta_back; eta_t = eta_front; normal = -V3{0, 0, 1}; }
    V3 t;
    if (!refract(normal, wo, eta_i / eta_t, &t)) {
        *wi = V3{0, 0, 0};
        *c = black();
        return;
    }
    float f_tr = 1.0f - fresnel_refl_coeff(l.fresnel, cos_theta(t));
    *wi = t;
    *c = (f_tr / std::fabs(cos_theta(t))) * l.albedo;
}

Color lobe_eval(const Lobe &l, V3 wo, V3 wi);
Prob lobe_prob(const Lobe &l, V3 wo, V3 wi);

// bxdf.rs:462-501, 560-564, 611-626
inline void lobe_sample(const Lobe &l, V3 wo, float r0, float r1, Color *f, V3 *wi, Prob *pr) {
    switch (l.kind) {
    case LOBE_SPECULAR: {
        if (l.intrusion == INTR_REFLECTION) {
            specular_reflect(l, wo, wi, f);
            *pr = Mass(1.0f);
        } else if (l.intrusion == INTR_TRANSMISSION) {
            specular_refract(l, wo, l.fresnel.eta_front, l.fresnel.eta_back, wi, f);
            *pr = Mass(1.0f);
        } else {
            float rc = fresnel_refl_coeff(l.fresnel, cos_theta(wo));
            if (r0 < rc) {
                specular_reflect(l, wo, wi, f);
                *pr = Mass(rc);
            } else {
                specular_refract(l, wo, l.fresnel.eta_front, l.fresnel.eta_back, wi, f);
                *pr = Mass(1.0f - rc);
            }
        }
        return;
    }
    case LOBE_LAMBERT:
    case LOBE_OREN_NAYAR: {
        if (!(cos_theta(wo) >= 0.0f)) panic_flag(P_LAMBERT_WO);
        *wi = cos_sample_hemisphere(r0, r1);
        *f = lobe_eval(l, wo, *wi);
        *pr = lobe_prob(l, wo, *wi);
        return;
    }
    default: {  // LOBE_MICROFACET
        V3 wh = mf_sample_wh(l.distrib, wo, r0, r1);
        V3 w = reflect(wh, wo);
        if (!same_hemisphere(wo, w)) {
            *f = black();
            *wi = V3{0, 0, 1};
            *pr = Density(0.0f);
            return;
        }
        float pdf = mf_pdf(l.distrib, wo, wh) / (4.0f * dot(wo, wh));
        *f = lobe_eval(l, wo, w);
        *wi = w;
        *pr = Density(pdf);
        return;
    }
    }
}
// bxdf.rs:458-460, 540-559, 594-609
inline Color lobe_eval(const Lobe &l, V3 wo, V3 wi) {
    switch (l.kind) {
    case LOBE_SPECULAR: return black();
    case LOBE_LAMBERT: return l.albedo * kFrac1Pi;
    case LOBE_OREN_NAYAR: {
        float sti = sin_theta(wi), sto = sin_theta(wo);
        float spi, cpi, spo, cpo;
        sin_cos_phi(wi, &spi, &cpi);
        sin_cos_phi(wo, &spo, &cpo);
        float dcp = f_max(cpi * cpo + spi * spo, 0.0f);
        float aci = std::fabs(cos_theta(wi)), aco = std::fabs(cos_theta(wo));
        float sin_alpha, tan_beta;
        if (aci > aco) { sin_alpha = sto; tan_beta = sti / aci; }
        else { sin_alpha = sti; tan_beta = sto / aco; }
        return l.albedo * kFrac1Pi * (l.on_a + l.on_b * dcp * sin_alpha * tan_beta);
    }
    default: {
        float cto = std::fabs(cos_theta(wo));
        float cti = std::fabs(cos_theta(wi));
        V3 wh;
        bool ok = bisector(wo, wi, &wh);
        if (cto == 0.0f || cti == 0.0f || !ok) return black();
        wh = face_forward(wh, V3{0, 0, 1});
        Color refl = fresnel_eval(l.fresnel, dot(wi, wh));
        return l.albedo * mf_d(l.distrib, wh) * mf_g(l.distrib, wo, wi) * refl *
               weak_recip(4.0f * cto * cti);
    }
    }
}
// bxdf.rs:503-505, 566-572, 628-638
inline Prob lobe_prob(const Lobe &l, V3 wo, V3 wi) {
    switch (l.kind) {
    case LOBE_SPECULAR: return Mass(0.0f);
    case LOBE_LAMBERT:
    case LOBE_OREN_NAYAR:
        if (wo.z * wi.z >= 0.0f) return Density(cos_hemisphere_pdf(wi));
        return Density(0.0f);
    default: {
        if (!same_hemisphere(wo, wi)) return Density(0.0f);
        V3 wh;
        if (bisector(wo, wi, &wh)) return Density(mf_pdf(l.distrib, wo, wh) / (4.0f * dot(wo, wh)));
        return Density(0.0f);
    }
    }
}

}  // namespace orc
