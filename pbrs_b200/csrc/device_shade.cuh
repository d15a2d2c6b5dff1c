// Shading arithmetic: local-frame trigonometry, Fresnel, Beckmann microfacets, the BxDF lobes,
// textures, material -> lobe construction, the multi-lobe BSDF, lights and the environment.
//
// Same operation order as the reference functions cited on each block (see device_math.cuh for
// the parity contract).  Lobes live in a fixed array of at most five 64-byte records per hit
// instead of the reference's heap-allocated Vec<BXDF> (material/src/lib.rs:180-445 builds one per
// call, and the integrators call it two or three times per bounce; here it is built once).
#pragma once
#include "device_geom.cuh"
#include "device_simple.cuh"

namespace pbrs {

// ---- Omega: geometry/src/bxdf.rs:42-177 (local shading frame, +Z = normal) ----
PB_DEV float cos_theta(vec3 w) { return w.z; }
PB_DEV float cos2_theta(vec3 w) { return w.z * w.z; }
PB_DEV float sin2_theta(vec3 w) { return 1.0f - cos2_theta(w); }
PB_DEV float sin_theta(vec3 w) { return sqrtf(fmaxf(sin2_theta(w), 0.0f)); }
PB_DEV float tan2_theta(vec3 w) { return sin2_theta(w) / cos2_theta(w); }
PB_DEV float div_or(float x, float d, float fallback) { return d == 0.0f ? fallback : x / d; }
PB_DEV float cos2_phi(vec3 w) { return div_or(w.x * w.x, w.x * w.x + w.y * w.y, 1.0f); }
PB_DEV float sin2_phi(vec3 w) { return div_or(w.y * w.y, w.x * w.x + w.y * w.y, 0.0f); }
PB_DEV bool same_hemisphere(vec3 a, vec3 b) { return a.z * b.z >= 0.0f; }     // :111-113
PB_DEV bool bisector(vec3 a, vec3 b, vec3 &out) { return try_hat(a + b, out); }  // :143-146
PB_DEV vec3 face_forward(vec3 w, vec3 n) { return sign_neg(dot(w, n)) ? -w : w; }  // :149-155

// bxdf.rs:187-206
PB_DEV void concentric_sample_disk(float u, float v, float &ox, float &oy) {
    float x = u * 2.0f - 1.0f;
    float y = v * 2.0f - 1.0f;
    if (x == 0.0f && y == 0.0f) { ox = 0.0f; oy = 0.0f; return; }
    float r = fabsf(fabsf(x) > fabsf(y) ? x : y);
    float hyp = t_hypot(x, y);
    float ct = x / hyp, st = y / hyp;
    ox = r * ct;
    oy = r * st;
}
PB_DEV vec3 cos_sample_hemisphere(float u, float v) {
    float x, y;
    concentric_sample_disk(u, v, x, y);
    float z = sqrtf(fmaxf(1.0f - x * x - y * y, 0.0f));
    return mk(x, y, z);
}

#ifndef PBRS_INLINE_TEXENV
#define PBRS_INLINE_TEXENV 1  // texture_value and eval_env inlined: shade -1.8 % on C4, -1.3 % on C3, +0.7 % on C5 (part 8 of the same log)
#endif
#if PBRS_INLINE_TEXENV
#define PB_CALL_TEXENV PB_DEV
#else
#define PB_CALL_TEXENV PB_CALL
#endif
#ifndef PBRS_INLINE_DYN
#define PBRS_INLINE_DYN 1  // the multi-lobe class: its dynamic BSDF and light-sampling routines inlined (shade -2.5 % on C4, profiles/r2_exp_shade_inline_gridconstant.log part 7)
#endif
#if PBRS_INLINE_DYN
#define PB_CALL_DYN PB_DEV
#else
#define PB_CALL_DYN PB_CALL
#endif
// ---- lobes ----
enum { LOBE_SPECULAR = 0, LOBE_LAMBERT = 1, LOBE_MICROFACET = 2 };
enum { FR_NOP = 0, FR_DIELECTRIC = 1, FR_CONDUCTOR = 2 };
enum { INTR_REFLECTION = 0, INTR_TRANSMISSION = 1, INTR_HYBRID = 2 };
struct Lobe {
    int kind, fresnel, intrusion;
    color albedo;
    float eta_front, eta_back;  // dielectric Fresnel
    color eta_t, k;             // conductor Fresnel (eta_i = 1, bxdf.rs:299-305)
    float ax, ay;               // Beckmann alphas
};
#define PBRS_MAX_LOBES 5
struct Lobes {
    int n;
    Lobe l[PBRS_MAX_LOBES];
};

// ---- Fresnel: bxdf.rs:308-392 ----
PB_DEV float fresnel_refl_coeff(const Lobe &f, float cos_i, Diag &dg) {
    if (f.fresnel == FR_NOP) return 1.0f;
    if (f.fresnel == FR_CONDUCTOR) { flag(dg, P_FRESNEL); return 1.0f; }
    cos_i = clampf(cos_i, -1.0f, 1.0f);
    float eta_i, eta_t;
    if (cos_i > 0.0f) { eta_i = f.eta_front; eta_t = f.eta_back; }
    else { eta_i = f.eta_back; eta_t = f.eta_front; cos_i = -cos_i; }
    float sin_i = sqrtf(fmaxf(1.0f - cos_i * cos_i, 0.0f));
    float sin_t = eta_i / eta_t * sin_i;
    if (sin_t >= 1.0f) return 1.0f;
    float cos_t = sqrtf(fmaxf(1.0f - sin_t * sin_t, 0.0f));
    float r_perp = (eta_i * cos_i - eta_t * cos_t) / (eta_i * cos_i + eta_t * cos_t);
    float r_par = (eta_t * cos_i - eta_i * cos_t) / (eta_t * cos_i + eta_i * cos_t);
    return (r_par * r_par + r_perp * r_perp) * 0.5f;
}
PB_CALL color fresnel_eval(const Lobe &f, float cos_i, Diag &dg) {
    if (f.fresnel != FR_CONDUCTOR) return grayc(fresnel_refl_coeff(f, cos_i, dg));
    color eta_i = grayc(1.0f);
    color eta = cw_div(f.eta_t, eta_i);
    color eta2 = eta * eta;
    color etak = cw_div(f.k, eta_i);
    color etak2 = etak * etak;
    float c = clampf(cos_i, -1.0f, 1.0f);
    float cos2 = c * c;
    float sin2 = 1.0f - cos2;
    color t0 = eta2 - etak2 - grayc(sin2);
    color a2_plus_b2 = cw_sqrt(t0 * t0 + 4.0f * eta2 * etak2);
    color t1 = a2_plus_b2 + grayc(cos2);
    color a = cw_sqrt((a2_plus_b2 + t0) * 0.5f);
    color t2 = 2.0f * a * cos_i;
    color ratio_s = cw_div(t1 - t2, t1 + t2);
    if (!is_finite(ratio_s)) flag(dg, P_FRESNEL);
    color t3 = cos2 * a2_plus_b2 + grayc(sin2 * sin2);
    color t4 = t2 * sin2;
    color ratio_p = ratio_s * cw_div(t3 - t4, t3 + t4);
    if (!is_finite(ratio_p)) flag(dg, P_FRESNEL);
    return cw_max((ratio_s + ratio_p) * 0.5f, 0.0f);
}

// ---- Beckmann: geometry/src/microfacet.rs ----
// :16-23, left to right (Q5)
PB_DEV float roughness_to_alpha(float roughness) {
    float x = fmaxf(t_log(roughness), -8.0f);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
// :36-59
PB_CALL float mf_d(float ax, float ay, vec3 wh, Diag &dg) {
    float tan2 = tan2_theta(wh);
    float c2 = cos2_theta(wh);
    float cos4 = c2 * c2;
    if (is_nan(tan2) || is_nan(cos4)) flag(dg, P_MISC);
    if (is_inf(tan2)) return 0.0f;
    float x = cos2_phi(wh) / (ax * ax) + sin2_phi(wh) / (ay * ay);
    return t_exp(x * -tan2) / (kPi * ax * ay * cos4);
}
// :64-88
PB_CALL float mf_lambda(float ax, float ay, vec3 w) {
    float abs_tan = fabsf(sqrtf(tan2_theta(w)));
    if (is_inf(abs_tan)) return 0.0f;
    float alpha = sqrtf(cos2_phi(w) * (ax * ax) + sin2_phi(w) * (ay * ay));
    float a = 1.0f / (alpha * abs_tan);
    if (a >= 1.6f) return 0.0f;
    return (1.0f - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
}
// :107-109
PB_DEV float mf_g(float ax, float ay, vec3 wo, vec3 wi) { return 1.0f / (1.0f + mf_lambda(ax, ay, wo) + mf_lambda(ax, ay, wi)); }
// :111-124, the cfg(not(sample_visible_area)) branch
PB_DEV float mf_pdf(float ax, float ay, vec3 wh, Diag &dg) {
    float d = mf_d(ax, ay, wh, dg);
    float y = fabsf(cos_theta(wh));
    if (is_nan(d * y)) flag(dg, P_MISC);
    return d * y;
}
// :126-159
PB_CALL vec3 mf_sample_wh(float ax, float ay, vec3 wo, float u, float v, Diag &dg) {
    float tan2, phi;
    float log_sample = t_log(1.0f - u);
    if (!is_fin(log_sample)) flag(dg, P_LOG_SAMPLE);
    if (ax == ay) {
        tan2 = -(ax * ax) * log_sample;
        phi = v * 2.0f * kPi;
    } else {
        phi = t_atan(ay / ax * t_tan(2.0f * kPi * v + kHalfPi));
        if (v >= 0.5f) phi += kPi;
        float sp = t_sin(phi), cp = t_cos(phi);
        float ca = cp / ax, sa = sp / ay;
        float alpha2 = ca * ca + sa * sa;
        tan2 = -log_sample / alpha2;
    }
    float ct = 1.0f / sqrtf(1.0f + tan2);
    float st = ct * sqrtf(tan2);
    vec3 wh = spherical_direction(st, ct, phi);
    return face_forward(wh, wo);
}

struct Prob {
    bool is_mass;
    float v;
};
PB_DEV Prob Mass(float m) { Prob p; p.is_mass = true; p.v = m; return p; }
PB_DEV Prob Density(float d) { Prob p; p.is_mass = false; p.v = d; return p; }

// ---- BxDF impls: bxdf.rs:427-639 ----
PB_DEV void specular_reflect(const Lobe &l, vec3 wo, vec3 &wi, color &c, Diag &dg) {  // :427-434
    wi = mk(-wo.x, -wo.y, wo.z);
    color fr = fresnel_eval(l, cos_theta(wi), dg);
    c = fr * l.albedo * weak_recip(fabsf(cos_theta(wi)));
}
PB_DEV void specular_refract(const Lobe &l, vec3 wo, vec3 &wi, color &c, Diag &dg) {  // :436-454
    float eta_i, eta_t;
    vec3 normal;
    if (cos_theta(wo) > 0.0f) { eta_i = l.eta_front; eta_t = l.eta_back; normal = mk(0.0f, 0.0f, 1.0f); }
    else { eta_i = l.eta_back; eta_t = l.eta_front; normal = -mk(0.0f, 0.0f, 1.0f); }
    vec3 t;
    if (!refract(normal, wo, eta_i / eta_t, t, dg)) { wi = mk(0.0f, 0.0f, 0.0f); c = blackc(); return; }
    float f_tr = 1.0f - fresnel_refl_coeff(l, cos_theta(t), dg);
    wi = t;
    c = (f_tr / fabsf(cos_theta(t))) * l.albedo;
}
// K >= 0: the lobe kind is a compile-time constant (material-class kernels); K < 0: read l.kind.
// The *_t templates hold the arithmetic; the un-suffixed entry points inline them for a known K
// and go through one out-of-line generic copy (*_dyn) otherwise.
template <int K>
PB_DEV int kind_of(const Lobe &l) { return K >= 0 ? K : l.kind; }

// :458-460, 540-559 (Lambert only: Oren-Nayar is never instantiated by a material), 594-609
template <int K>
PB_DEV color lobe_eval_t(const Lobe &l, vec3 wo, vec3 wi, Diag &dg) {
    const int kind = kind_of<K>(l);
    if (kind == LOBE_SPECULAR) return blackc();
    if (kind == LOBE_LAMBERT) return l.albedo * kInvPi;
    float cto = fabsf(cos_theta(wo));
    float cti = fabsf(cos_theta(wi));
    vec3 wh;
    bool ok = bisector(wo, wi, wh);
    if (cto == 0.0f || cti == 0.0f || !ok) return blackc();
    wh = face_forward(wh, mk(0.0f, 0.0f, 1.0f));
    color refl = fresnel_eval(l, dot(wi, wh), dg);
    return l.albedo * mf_d(l.ax, l.ay, wh, dg) * mf_g(l.ax, l.ay, wo, wi) * refl * weak_recip(4.0f * cto * cti);
}
// :503-505, 566-572 (Q6), 628-638
template <int K>
PB_DEV Prob lobe_prob_t(const Lobe &l, vec3 wo, vec3 wi, Diag &dg) {
    const int kind = kind_of<K>(l);
    if (kind == LOBE_SPECULAR) return Mass(0.0f);
    if (kind == LOBE_LAMBERT) {
        if (wo.z * wi.z >= 0.0f) return Density(wi.z * kInvPi);
        return Density(0.0f);
    }
    if (!same_hemisphere(wo, wi)) return Density(0.0f);
    vec3 wh;
    if (bisector(wo, wi, wh)) return Density(mf_pdf(l.ax, l.ay, wh, dg) / (4.0f * dot(wo, wh)));
    return Density(0.0f);
}
// :462-501, 560-564, 611-626
template <int K>
PB_DEV void lobe_sample_t(const Lobe &l, vec3 wo, float r0, float r1, color &f, vec3 &wi, Prob &pr, Diag &dg) {
    const int kind = kind_of<K>(l);
    if (kind == LOBE_SPECULAR) {
        if (l.intrusion == INTR_REFLECTION) {
            specular_reflect(l, wo, wi, f, dg);
            pr = Mass(1.0f);
        } else if (l.intrusion == INTR_TRANSMISSION) {
            specular_refract(l, wo, wi, f, dg);
            pr = Mass(1.0f);
        } else {
            float rc = fresnel_refl_coeff(l, cos_theta(wo), dg);
            if (r0 < rc) { specular_reflect(l, wo, wi, f, dg); pr = Mass(rc); }
            else { specular_refract(l, wo, wi, f, dg); pr = Mass(1.0f - rc); }
        }
        return;
    }
    if (kind == LOBE_LAMBERT) {
        if (!(cos_theta(wo) >= 0.0f)) flag(dg, P_LAMBERT_WO);
        wi = cos_sample_hemisphere(r0, r1);
        f = lobe_eval_t<LOBE_LAMBERT>(l, wo, wi, dg);
        pr = lobe_prob_t<LOBE_LAMBERT>(l, wo, wi, dg);
        return;
    }
    vec3 wh = mf_sample_wh(l.ax, l.ay, wo, r0, r1, dg);
    vec3 w = reflect(wh, wo);
    if (!same_hemisphere(wo, w)) { f = blackc(); wi = mk(0.0f, 0.0f, 1.0f); pr = Density(0.0f); return; }
    float pdf = mf_pdf(l.ax, l.ay, wh, dg) / (4.0f * dot(wo, wh));
    f = lobe_eval_t<LOBE_MICROFACET>(l, wo, w, dg);
    wi = w;
    pr = Density(pdf);
}
PB_CALL color lobe_eval_dyn(const Lobe &l, vec3 wo, vec3 wi, Diag &dg) { return lobe_eval_t<-1>(l, wo, wi, dg); }
PB_CALL Prob lobe_prob_dyn(const Lobe &l, vec3 wo, vec3 wi, Diag &dg) { return lobe_prob_t<-1>(l, wo, wi, dg); }
PB_CALL void lobe_sample_dyn(const Lobe &l, vec3 wo, float r0, float r1, color &f, vec3 &wi, Prob &pr, Diag &dg) {
    lobe_sample_t<-1>(l, wo, r0, r1, f, wi, pr, dg);
}
template <int K>
PB_DEV color lobe_eval(const Lobe &l, vec3 wo, vec3 wi, Diag &dg) {
    if constexpr (K < 0) return lobe_eval_dyn(l, wo, wi, dg);
    else return lobe_eval_t<K>(l, wo, wi, dg);
}
template <int K>
PB_DEV Prob lobe_prob(const Lobe &l, vec3 wo, vec3 wi, Diag &dg) {
    if constexpr (K < 0) return lobe_prob_dyn(l, wo, wi, dg);
    else return lobe_prob_t<K>(l, wo, wi, dg);
}
template <int K>
PB_DEV void lobe_sample(const Lobe &l, vec3 wo, float r0, float r1, color &f, vec3 &wi, Prob &pr, Diag &dg) {
    if constexpr (K < 0) lobe_sample_dyn(l, wo, r0, r1, f, wi, pr, dg);
    else lobe_sample_t<K>(l, wo, r0, r1, f, wi, pr, dg);
}

// ---- textures: texture/src/lib.rs ----
PB_DEV color unpack_rgb8(uint32_t p) {  // Color::rgb(u8,u8,u8), radiometry/src/color.rs:49-51
    return mkc((float)(p & 255u) / 255.0f, (float)((p >> 8) & 255u) / 255.0f, (float)((p >> 16) & 255u) / 255.0f);
}
// lib.rs:98-138
PB_CALL float perlin_noise(const DeviceScene &sc, const TextureRec &t, vec3 p, Diag &dg) {
    float fx = p.x * t.freq, fy = p.y * t.freq, fz = p.z * t.freq;
    float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    int i = (int)flx, j = (int)fly, k = (int)flz;
    float u = fx - flx, v = fy - fly, w = fz - flz;
    u = u * u * (3.0f - 2.0f * u);
    v = v * v * (3.0f - 2.0f * v);
    w = w * w * (3.0f - 2.0f * w);
    const float *rv = sc.perlin_vec + 768u * t.perlin_base;
    const uint32_t *pm = sc.perlin_perm + 768u * t.perlin_base;
    float accum = 0.0f;
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk) {
                uint32_t ii = (uint32_t)((i + di) & 255), jj = (uint32_t)((j + dj) & 255), kk = (uint32_t)((k + dk) & 255);
                uint32_t index = ld_u32(pm + ii) ^ ld_u32(pm + 256 + jj) ^ ld_u32(pm + 512 + kk);
                vec3 c = mk(ld_f32(rv + 3 * index), ld_f32(rv + 3 * index + 1), ld_f32(rv + 3 * index + 2));
                vec3 wv = mk(u - (float)di, v - (float)dj, w - (float)dk);
                float dp = dot(c, wv);
                accum += ((float)di * u + (float)(1 - di) * (1.0f - u)) * ((float)dj * v + (float)(1 - dj) * (1.0f - v)) *
                         ((float)dk * w + (float)(1 - dk) * (1.0f - w)) * dp;
            }
    if (!(accum >= -1.0f) || !(accum <= 1.0f)) flag(dg, P_PERLIN);
    return accum;
}
// lib.rs:140-147
PB_DEV float perlin_turbulence(const DeviceScene &sc, const TextureRec &t, vec3 p, Diag &dg) {
    float acc = 0.0f;
    float s = 1.0f, wgt = 1.0f;  // 2^i and 0.5^i are exact
    for (int i = 0; i < 7; ++i) {
        acc = acc + wgt * perlin_noise(sc, t, mk(p.x * s, p.y * s, p.z * s), dg);
        s *= 2.0f;
        wgt *= 0.5f;
    }
    return fabsf(acc);
}
PB_DEV color image_lookup(const DeviceScene &sc, const TextureRec &t, float u, float v) {  // lib.rs:211-223
    u = clampf(u, 0.0f, 1.0f);
    v = clampf(v, 0.0f, 1.0f);
    float fu = u * (float)t.width, fv = v * (float)t.height;
    // `as usize` saturates: NaN / negative -> 0
    uint32_t col = (fu > 0.0f ? (uint32_t)fu : 0u) % t.width;
    uint32_t row = (fv > 0.0f ? (uint32_t)fv : 0u) % t.height;
    return unpack_rgb8(ld_u32(sc.texels + t.texel_base + row * t.width + col));
}
PB_CALL_TEXENV color texture_value(const DeviceScene &sc, int id, float u, float v, vec3 p, Diag &dg) {
    const TextureRec &t = sc.textures[id];
    if (t.kind == PBRS_TEX_SOLID) return mkc(t.value[0], t.value[1], t.value[2]);  // :29-33
    if (t.kind == PBRS_TEX_IMAGE) return image_lookup(sc, t, u, v);
    // :150-160 marble: sin(...).mul_add(0.5, 0.5) * white
    float s = fmaf(t_sin(t.freq * p.z + 10.0f * perlin_turbulence(sc, t, p, dg)), 0.5f, 0.5f);
    return s * grayc(1.0f);
}

// ---- materials -> lobes: material/src/lib.rs ----
PB_DEV Lobe mk_lambert(color albedo) {
    Lobe l;
    l.kind = LOBE_LAMBERT; l.fresnel = FR_NOP; l.intrusion = 0; l.albedo = albedo;
    l.eta_front = l.eta_back = 0.0f; l.eta_t = blackc(); l.k = blackc(); l.ax = l.ay = 0.0f;
    return l;
}
PB_DEV Lobe mk_specular(color albedo, int intrusion, int fresnel, float eta_o, float eta_i) {
    Lobe l = mk_lambert(albedo);
    l.kind = LOBE_SPECULAR; l.intrusion = intrusion; l.fresnel = fresnel; l.eta_front = eta_o; l.eta_back = eta_i;
    return l;
}
PB_DEV Lobe mk_microfacet(color albedo, float ax, float ay) {
    Lobe l = mk_lambert(albedo);
    l.kind = LOBE_MICROFACET; l.ax = ax; l.ay = ay;
    return l;
}
PB_DEV color mtl_emission(const MaterialRec &m) {  // :291-299
    return m.kind == PBRS_MTL_DIFFUSE_LIGHT ? mkc(m.a[0], m.a[1], m.a[2]) : blackc();
}
// CLS: the material class the caller is specialised for (PBRS_CLS_ANY: none); only the material
// kinds of that class are compiled in.
PB_DEV constexpr bool cls_has(int cls, int kind) {
    return cls == PBRS_CLS_ANY ||
           (cls == PBRS_CLS_LAMBERT && (kind == PBRS_MTL_LAMBERTIAN || kind == PBRS_MTL_SUBSTRATE)) ||
           (cls == PBRS_CLS_MICROFACET && (kind == PBRS_MTL_METAL || kind == PBRS_MTL_GLOSSY)) ||
           (cls == PBRS_CLS_SPECULAR && (kind == PBRS_MTL_MIRROR || kind == PBRS_MTL_DIELECTRIC)) ||
           (cls == PBRS_CLS_MULTI && (kind == PBRS_MTL_PLASTIC || kind == PBRS_MTL_UBER));
}
// the lobe kind every lobe of a class has, or -1
PB_DEV constexpr int cls_lobe_kind(int cls) {
    return cls == PBRS_CLS_LAMBERT ? LOBE_LAMBERT : cls == PBRS_CLS_MICROFACET ? LOBE_MICROFACET : cls == PBRS_CLS_SPECULAR ? LOBE_SPECULAR
           : cls == PBRS_CLS_EMISSIVE ? LOBE_LAMBERT /* no lobes at all */ : -1;
}
template <int CLS>
PB_DEV void bxdfs_at_t(const DeviceScene &sc, const MaterialRec &m, const Isect &h, Lobes &L, Diag &dg) {
    L.n = 0;
    color ca = mkc(m.a[0], m.a[1], m.a[2]), cb = mkc(m.b[0], m.b[1], m.b[2]);
    // A class whose lobes are all of one kind (Lambert, microfacet, specular) has materials of at most
    // ONE lobe: it goes to slot 0 with a static index, and the BSDF routines below read slot 0 only, so
    // the lobe stays in registers (a dynamically indexed Lobes lives in local memory: the Lambert scatter
    // kernel read its albedo back from there ~25 times per path).
    constexpr bool single = cls_lobe_kind(CLS) >= 0;
    auto put = [&L](const Lobe &l) {
        if constexpr (single) { L.l[0] = l; L.n = 1; }
        else L.l[L.n++] = l;
    };
    if (cls_has(CLS, PBRS_MTL_LAMBERTIAN) && m.kind == PBRS_MTL_LAMBERTIAN) {  // :180-184
        put(mk_lambert(texture_value(sc, m.tex_kd, h.u, h.v, h.pos, dg)));
    }
    if (cls_has(CLS, PBRS_MTL_METAL) && m.kind == PBRS_MTL_METAL) {  // :200-206
        float alpha = roughness_to_alpha(m.f[0]);
        Lobe l = mk_microfacet(grayc(1.0f), alpha, alpha);
        l.fresnel = FR_CONDUCTOR; l.eta_t = ca; l.k = cb;
        put(l);
    }
    if (cls_has(CLS, PBRS_MTL_GLOSSY) && m.kind == PBRS_MTL_GLOSSY) {  // :72-78, :216-218
        float alpha = roughness_to_alpha(m.f[0]);
        put(mk_microfacet(ca, alpha, alpha));
    }
    if (cls_has(CLS, PBRS_MTL_MIRROR) && m.kind == PBRS_MTL_MIRROR) {  // :229-232
        put(mk_specular(ca, INTR_REFLECTION, FR_NOP, 0.0f, 0.0f));
    }
    if (cls_has(CLS, PBRS_MTL_DIELECTRIC) && m.kind == PBRS_MTL_DIELECTRIC) {  // :265-268
        put(mk_specular(ca, INTR_HYBRID, FR_DIELECTRIC, 1.0f, m.f[0]));
    }
    // DiffuseLight (:291-293): no lobes
    if (cls_has(CLS, PBRS_MTL_PLASTIC) && m.kind == PBRS_MTL_PLASTIC) {  // :433-445
        float alpha = m.remap ? roughness_to_alpha(m.f[0]) : m.f[0];
        L.l[L.n++] = mk_microfacet(cb, alpha, alpha);
        L.l[L.n++] = mk_lambert(ca);
    }
    if (cls_has(CLS, PBRS_MTL_UBER) && m.kind == PBRS_MTL_UBER) {  // :317-365
        color transmission = grayc(clampf(1.0f - m.f[3], 0.0f, 1.0f));
        if (!is_black(transmission)) L.l[L.n++] = mk_specular(transmission, INTR_TRANSMISSION, FR_DIELECTRIC, 1.0f, m.f[2]);
        color kd = texture_value(sc, m.tex_kd, h.u, h.v, h.pos, dg);
        if (!is_black(kd)) L.l[L.n++] = mk_lambert(kd);
        color ks = texture_value(sc, m.tex_ks, h.u, h.v, h.pos, dg);
        if (!is_black(ks)) {
            float au = m.remap ? roughness_to_alpha(m.f[0]) : m.f[0];
            float av = m.remap ? roughness_to_alpha(m.f[1]) : m.f[1];
            Lobe l = mk_microfacet(ks, au, av);
            l.fresnel = FR_DIELECTRIC; l.eta_front = 1.0f; l.eta_back = m.f[2];
            L.l[L.n++] = l;
        }
        if (m.tex_kr >= 0) {
            color kr = texture_value(sc, m.tex_kr, h.u, h.v, h.pos, dg);
            if (!is_black(kr)) L.l[L.n++] = mk_specular(kr, INTR_HYBRID, FR_DIELECTRIC, 1.0f, m.f[2]);
        }
        if (m.tex_kt >= 0) {
            color kt = texture_value(sc, m.tex_kt, h.u, h.v, h.pos, dg);
            if (!is_black(kt)) L.l[L.n++] = mk_specular(kt, INTR_TRANSMISSION, FR_DIELECTRIC, 1.0f, m.f[2]);
        }
    }
    if (cls_has(CLS, PBRS_MTL_SUBSTRATE) && m.kind == PBRS_MTL_SUBSTRATE) {  // :393-420 (FresnelBlend is commented out upstream: Lambert only)
        color d = texture_value(sc, m.tex_kd, h.u, h.v, h.pos, dg), s = texture_value(sc, m.tex_ks, h.u, h.v, h.pos, dg);
        if (!(is_black(d) && is_black(s))) put(mk_lambert(d));
    }
}

PB_CALL_DYN void bxdfs_at_dyn(const DeviceScene &sc, const MaterialRec &m, const Isect &h, Lobes &L, Diag &dg) { bxdfs_at_t<PBRS_CLS_ANY>(sc, m, h, L, dg); }
template <int CLS>
PB_DEV void bxdfs_at(const DeviceScene &sc, const MaterialRec &m, const Isect &h, Lobes &L, Diag &dg) {
    if constexpr (CLS == PBRS_CLS_ANY || CLS == PBRS_CLS_MULTI) bxdfs_at_dyn(sc, m, h, L, dg);
    else bxdfs_at_t<CLS>(sc, m, h, L, dg);
}

// ---- BSDF: src/bsdf.rs ----
struct Frame {
    vec3 t, b, n;
};
PB_DEV Frame bsdf_frame(const Isect &h, Diag &dg) {  // :18-31, :125-137
    Frame f;
    f.n = hat(h.normal, dg);
    f.b = hat(cross(h.normal, h.tangent), dg);
    f.t = cross(f.b, f.n);
    if (!(fabsf(dot(f.n, f.b)) < 1e-4f) || !(fabsf(dot(f.n, f.t)) < 1e-4f) || !(fabsf(dot(f.t, f.b)) < 1e-4f))
        flag(dg, P_BSDF_FRAME);
    float det = dot(cross(f.t, f.b), f.n);
    if (!(fabsf(det - 1.0f) < 1e-4f)) flag(dg, P_BSDF_FRAME);
    return f;
}
PB_DEV vec3 to_local(const Frame &f, vec3 w, Diag &dg) { return hat(mk(dot(f.t, w), dot(f.b, w), dot(f.n, w)), dg); }  // :113-117
PB_DEV vec3 to_world(const Frame &f, vec3 l) { return l.x * f.t + l.y * f.b + l.z * f.n; }                           // :119-123
template <int K>
PB_DEV color bsdf_eval_t(const Frame &fr, const Lobes &L, vec3 wo_w, vec3 wi_w, Diag &dg) {  // :43-51
    vec3 wi = to_local(fr, wi_w, dg);
    vec3 wo = to_local(fr, wo_w, dg);
    if (wo.z == 0.0f) return blackc();
    color s = blackc();
    if constexpr (K >= 0) { if (L.n > 0) s = s + lobe_eval<K>(L.l[0], wo, wi, dg); }  // single-lobe class: slot 0, static index
    else for (int i = 0; i < L.n; ++i) s = s + lobe_eval<K>(L.l[i], wo, wi, dg);
    return s;
}
template <int K>
PB_DEV float bsdf_pdf_t(const Frame &fr, const Lobes &L, vec3 wo_w, vec3 wi_w, Diag &dg) {  // :53-57 (Q4: a sum)
    vec3 wi = to_local(fr, wi_w, dg);
    vec3 wo = to_local(fr, wo_w, dg);
    float s = 0.0f;
    if constexpr (K >= 0) {
        if (L.n > 0) {
            Prob p = lobe_prob<K>(L.l[0], wo, wi, dg);
            s += p.is_mass ? 0.0f : p.v;
        }
    } else {
        for (int i = 0; i < L.n; ++i) {
            Prob p = lobe_prob<K>(L.l[i], wo, wi, dg);
            s += p.is_mass ? 0.0f : p.v;
        }
    }
    return s;
}
// :59-103.  The chosen lobe is swap_remove()d from the list: the rest are visited with the last
// lobe moved into the chosen slot.
template <int K>
PB_DEV void bsdf_sample_t(const Frame &fr, const Lobes &L, vec3 wo_world, float u, float v, color &f, vec3 &wi_out, Prob &pr,
                          Diag &dg) {
    if (!(u < 1.0f)) flag(dg, P_MISC);
    vec3 wo = to_local(fr, wo_world, dg);
    int n = L.n;
    if (n == 0) { f = blackc(); wi_out = mk(0.0f, 0.0f, 0.0f); pr = Mass(0.0f); return; }
    if constexpr (K >= 0) n = 1;  // single-lobe class
    float un = u * (float)n;
    int chosen = (int)un;
    if (chosen >= n) chosen = n - 1;
    if constexpr (K >= 0) chosen = 0;  // (what the two lines above give for n = 1: a static index)
    float remapped_u = fractf(un);
    color value;
    vec3 wi;
    Prob prob;
    lobe_sample<K>(L.l[chosen], wo, v, remapped_u, value, wi, prob, dg);  // Q2: (v, remapped_u)
    if (prob.is_mass) { f = value; wi_out = to_world(fr, wi); pr = prob; return; }
    int count = 0;
    float other_sum = 0.0f;
    color others = blackc();
    // order after swap_remove: 0..chosen-1, then (last), then chosen+1..n-2
    for (int pass = 0; pass < 2; ++pass)
        for (int k = 0; k < n - 1; ++k) {
            int src = (k == chosen) ? n - 1 : k;
            if (pass == 0) {
                Prob p = lobe_prob<K>(L.l[src], wo, wi, dg);
                if (!p.is_mass) { count++; other_sum += p.v; }
            } else {
                others = others + lobe_eval<K>(L.l[src], wo, wi, dg);
            }
        }
    float overall = (prob.v + other_sum) / (float)(1 + count);
    f = value + others;
    wi_out = to_world(fr, wi);
    pr = Density(overall);
}
PB_CALL_DYN color bsdf_eval_dyn(const Frame &fr, const Lobes &L, vec3 wo_w, vec3 wi_w, Diag &dg) { return bsdf_eval_t<-1>(fr, L, wo_w, wi_w, dg); }
PB_CALL_DYN float bsdf_pdf_dyn(const Frame &fr, const Lobes &L, vec3 wo_w, vec3 wi_w, Diag &dg) { return bsdf_pdf_t<-1>(fr, L, wo_w, wi_w, dg); }
PB_CALL_DYN void bsdf_sample_dyn(const Frame &fr, const Lobes &L, vec3 wo_world, float u, float v, color &f, vec3 &wi_out, Prob &pr, Diag &dg) {
    bsdf_sample_t<-1>(fr, L, wo_world, u, v, f, wi_out, pr, dg);
}
template <int K>
PB_DEV color bsdf_eval(const Frame &fr, const Lobes &L, vec3 wo_w, vec3 wi_w, Diag &dg) {
    if constexpr (K < 0) return bsdf_eval_dyn(fr, L, wo_w, wi_w, dg);
    else return bsdf_eval_t<K>(fr, L, wo_w, wi_w, dg);
}
template <int K>
PB_DEV float bsdf_pdf(const Frame &fr, const Lobes &L, vec3 wo_w, vec3 wi_w, Diag &dg) {
    if constexpr (K < 0) return bsdf_pdf_dyn(fr, L, wo_w, wi_w, dg);
    else return bsdf_pdf_t<K>(fr, L, wo_w, wi_w, dg);
}
template <int K>
PB_DEV void bsdf_sample(const Frame &fr, const Lobes &L, vec3 wo_world, float u, float v, color &f, vec3 &wi_out, Prob &pr, Diag &dg) {
    if constexpr (K < 0) bsdf_sample_dyn(fr, L, wo_world, u, v, f, wi_out, pr, dg);
    else bsdf_sample_t<K>(fr, L, wo_world, u, v, f, wi_out, pr, dg);
}
// :104-112
template <int K>
PB_DEV bool bsdf_sample_specular(const Frame &fr, const Lobes &L, vec3 wo_world, color &f, vec3 &wi_out, Prob &pr, Diag &dg) {
    if constexpr (K >= 0 && K != LOBE_SPECULAR) return false;
    vec3 wo = to_local(fr, wo_world, dg);
    if constexpr (K >= 0) {  // the specular class: one lobe, slot 0
        if (L.n > 0) {
            vec3 wi;
            lobe_sample<K>(L.l[0], wo, 0.0f, 0.0f, f, wi, pr, dg);
            wi_out = to_world(fr, wi);
            return true;
        }
        return false;
    } else {
        for (int i = 0; i < L.n; ++i)
            if (kind_of<K>(L.l[i]) == LOBE_SPECULAR) {
                vec3 wi;
                lobe_sample<K>(L.l[i], wo, 0.0f, 0.0f, f, wi, pr, dg);
                wi_out = to_world(fr, wi);
                return true;
            }
        return false;
    }
}

// ---- environment: scene/src/lib.rs:96-117; scene/src/preset.rs:25-51 ----
PB_CALL_TEXENV color eval_env(const DeviceScene &sc, vec3 dir, Diag &dg) {
    if (sc.env_kind == PBRS_ENV_KIND_CONSTANT) return mkc(sc.env_color[0], sc.env_color[1], sc.env_color[2]);
    if (sc.env_kind == PBRS_ENV_KIND_IMAGE) {
        float phi = t_atan2(dir.z, dir.x);
        float u = fractf(phi * kInvPi * 0.5f + 1.0f);
        float cos_t = dir.y / len(dir);
        float v = t_acos(cos_t) / kPi;
        return image_lookup(sc, sc.env_image, u, v) * mkc(sc.env_scale[0], sc.env_scale[1], sc.env_scale[2]);
    }
    if (sc.env_fn == PBRS_ENV_BLUE_SKY) {
        float y = (hat(dir, dg).y + 1.0f) * 0.5f;
        return mkc(0.5f, 0.7f, 1.0f) * y + grayc(1.0f) * (1.0f - y);
    }
    if (sc.env_fn == PBRS_ENV_DARK_ROOM) {
        float y = (hat(dir, dg).y + 1.0f) * 0.5f;
        return grayc(0.1f) * y + grayc(0.1f) * (1.0f - y);
    }
    color horizon = mkc(245.0f / 255.0f, 174.0f / 255.0f, 82.0f / 255.0f);
    color dome = mkc(109.0f / 255.0f, 150.0f / 255.0f, 204.0f / 255.0f);
    float tilt = t_acos(hat(dir, dg).y);
    if (tilt > kPi * 0.25f) return dome;
    if (tilt > 0.0f) {
        float t = tilt / (kPi * 0.25f);
        return dome * t + horizon * (1.0f - t);
    }
    return grayc(0.2f);
}

// ---- area-light shapes: light/src/sample_shape.rs ----
PB_DEV Isect isect_rayless(vec3 pos, float u, float v, vec3 normal, Diag &dg) {
    return isect_new(pos, 0.0f, u, v, normal, mk(0.0f, 0.0f, 0.0f), dg);
}
// :184-195
PB_DEV void sphere_sample(vec3 c, float radius, float u, float v, vec3 &pos, vec3 &normal) {
    float theta = 2.0f * kPi * u;
    float phi = t_acos(2.0f * v - 1.0f);
    vec3 dir = mk(t_sin(phi) * t_cos(theta), t_sin(phi) * t_sin(theta), 2.0f * v - 1.0f);
    pos = c + radius * dir;
    normal = dir;
}
// :197-236
PB_CALL_AREA void sphere_sample_towards(vec3 c, float radius, vec3 target, float u, float v, vec3 &pos, vec3 &normal, Diag &dg) {
    vec3 wc = c - target;
    float r2 = radius * radius;
    if (len2(wc) < r2) { sphere_sample(c, radius, u, v, pos, normal); return; }
    float sin_theta_max_2 = r2 / len2(wc);
    float cos_theta_max = sqrtf(fmaxf(1.0f - sin_theta_max_2, 0.0f));
    float cos_t = (1.0f - u) + u * cos_theta_max;
    float sin_t2 = fmaxf(1.0f - cos_t * cos_t, 0.0f);
    float phi = v * 2.0f * kPi;
    float dc = len(wc);
    float ds = dc * cos_t - sqrtf(fmaxf(r2 - len2(wc) * sin_t2, 0.0f));
    float cos_alpha = (len2(wc) + r2 - ds * ds) / (2.0f * dc * radius);
    float sin_alpha = sqrtf(fmaxf(1.0f - cos_alpha * cos_alpha, 0.0f));
    vec3 n_obj = spherical_direction(sin_alpha, cos_alpha, phi);
    vec3 wcx, wcy;
    vec3 mz = -hat(wc, dg);
    make_coord_system(mz, wcx, wcy, dg);
    vec3 n_world = wcx * n_obj.x + wcy * n_obj.y + mz * n_obj.z;  // Mat3 * Vec3, hcm.rs:448-453
    pos = n_world * radius + c;
    normal = n_world;
}
PB_DEV float sphere_area(float radius) { return radius * radius * 4.0f * kPi; }  // :253-255
// :238-251
PB_DEV bool sphere_pdf_at(vec3 c, float radius, vec3 ref, vec3 wi, float &pdf) {
    vec3 rc = c - ref;
    float r2 = radius * radius;
    if (len2(rc) < r2) { pdf = 1.0f / sphere_area(radius); return true; }
    float sin_theta_max_2 = r2 / len2(rc);
    float cos_theta_max = sqrtf(fmaxf(1.0f - sin_theta_max_2, 0.0f));
    float cos_t = dot(rc, wi) / (len(rc) * len(wi));
    if (cos_t > cos_theta_max) { pdf = 1.0f / (2.0f * kPi * (1.0f - cos_theta_max)); return true; }
    return false;
}
// IsolatedTriangle::intersect, shape/src/simple.rs:425-427: only pos and normal are consumed
PB_DEV bool isotri_intersect(vec3 p0, vec3 p1, vec3 p2, const Ray &r, vec3 &pos, vec3 &normal, Diag &dg) {
    TriHit h;
    if (!tri_intersect(p0, p1, p2, r, h, dg)) return false;
    Isect i = isect_new(h.pos, h.t, h.b1, h.b2, h.normal, -r.d, dg);
    with_dpdu(i, p1 - p0, dg);
    pos = h.pos;
    normal = h.normal;
    return true;
}

struct AreaLight {
    int kind;
    vec3 p0, p1, p2;  // sphere: p0 = centre, p1.x = radius
    color emit;
    float area;
};
PB_DEV AreaLight load_area_light(const AreaLightRec *r) {
    AreaLight l;
    l.kind = r->kind;
    l.p0 = mk(r->p0[0], r->p0[1], r->p0[2]);
    l.p1 = mk(r->p1[0], r->p1[1], r->p1[2]);
    l.p2 = mk(r->p2[0], r->p2[1], r->p2[2]);
    l.emit = mkc(r->emit[0], r->emit[1], r->emit[2]);
    l.area = r->area;
    return l;
}
PB_CALL_AREA bool area_shape_intersect(const AreaLight &l, const Ray &r, vec3 &pos, vec3 &normal, Diag &dg) {
    if (l.kind == PBRS_AREA_SPHERE) {
        Isect h;
        if (!sphere_intersect(l.p0, l.p1.x, r, h, dg, false)) return false;
        pos = h.pos; normal = h.normal;
        return true;
    }
    if (l.kind == PBRS_AREA_TRIANGLE) return isotri_intersect(l.p0, l.p1, l.p2, r, pos, normal, dg);
    Isect h;
    if (l.kind == PBRS_AREA_QUAD) {
        if (!quad_intersect(l.p0, l.p1, l.p2, r, h, dg)) return false;
    } else {
        if (!disk_intersect(l.p0, l.p1, l.p2, r, h, dg)) return false;
    }
    pos = h.pos; normal = h.normal;
    return true;
}
// sample_shape.rs:28-33 default pdf_at (Q12: distance, not distance squared); sphere override :238
PB_CALL_AREA bool area_shape_pdf_at(const AreaLight &l, const Isect &ref, vec3 wi, float &pdf, Diag &dg) {
    if (l.kind == PBRS_AREA_SPHERE) return sphere_pdf_at(l.p0, l.p1.x, ref.pos, wi, pdf);
    Ray ray = spawn_ray(ref, wi);
    vec3 pos, normal;
    if (!area_shape_intersect(l, ray, pos, normal, dg)) return false;
    pdf = len(ref.pos - pos) / (fabsf(dot(normal, -wi)) * l.area);
    return true;
}
// sample_shape.rs:197 (sphere), :275-293 (triangle), :296-305 (quad), :257-269 (disk)
PB_DEV void area_shape_sample_towards(const AreaLight &l, const Isect &target, float u, float v, vec3 &pos, vec3 &normal,
                                      Diag &dg) {
    if (l.kind == PBRS_AREA_SPHERE) {
        sphere_sample_towards(l.p0, l.p1.x, target.pos, u, v, pos, normal, dg);
        (void)isect_rayless(pos, u, v, normal, dg);
        return;
    }
    if (l.kind == PBRS_AREA_QUAD) {  // :296-305: the normal is the unnormalised cross product
        pos = l.p0 + u * l.p1 + v * l.p2;
        normal = cross(l.p1, l.p2);
        return;
    }
    if (l.kind == PBRS_AREA_DISK) {  // :257-269
        float cos_t, sin_t;
        concentric_sample_disk(u, v, cos_t, sin_t);
        vec3 radial2 = cross(l.p1, l.p2);
        vec3 cp = l.p2 * cos_t + radial2 * sin_t;
        pos = l.p0 + cp;
        normal = facing(l.p1, target.normal);
        return;
    }
    if (u + v > 1.0f) { float nu = 1.0f - v, nv = 1.0f - u; u = nu; v = nv; }
    pos = l.p0 + (l.p1 - l.p0) * u + (l.p2 - l.p0) * v;
    normal = hat(cross(l.p0 - l.p1, l.p2 - l.p1), dg);
}

}  // namespace pbrs
