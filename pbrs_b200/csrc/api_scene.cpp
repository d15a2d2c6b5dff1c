// The scene-construction half of the C ABI of include/pbrs_gpu.h (the render half is
// api_render.cu).  Nothing here computes: it validates arguments the way the reference's
// constructors assert them, records the scene description, and at commit drives scene_host.cpp
// (BVH build) and scene_device.cu (upload).  Plain C++: no CUDA calls in this file.
#include <cmath>
#include <cstring>
#include <new>

#include "device_math.cuh"
#include "scene_host.h"

using namespace pbrs;

struct pbrs_scene {
    SceneImpl impl;
};

namespace {

int fail(int code, const char *msg) {
    set_error(msg);
    return code;
}
bool finite3(const float *v) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

#define NEED(cond, msg)                                      \
    do {                                                     \
        if (!(cond)) return fail(PBRS_ERR_INVALID_ARG, msg); \
    } while (0)
#define NOT_COMMITTED(s) \
    if ((s)->impl.committed) return fail(PBRS_ERR_STATE, "the scene is already committed")

HostTexture make_image(uint32_t w, uint32_t h, const uint8_t *rgb) {
    HostTexture t;
    std::memset(&t.rec, 0, sizeof t.rec);
    t.rec.kind = PBRS_TEX_IMAGE;
    t.rec.width = w;
    t.rec.height = h;
    t.texels.resize((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i)
        t.texels[i] = (uint32_t)rgb[3 * i] | ((uint32_t)rgb[3 * i + 1] << 8) | ((uint32_t)rgb[3 * i + 2] << 16) | 0xFF000000u;
    return t;
}

}  // namespace

extern "C" {

int pbrs_abi_version(void) { return PBRS_ABI_VERSION; }
const char *pbrs_last_error(void) { return get_error(); }

pbrs_scene *pbrs_scene_create(void) { return new (std::nothrow) pbrs_scene(); }
void pbrs_scene_destroy(pbrs_scene *s) {
    if (!s) return;
    device_free(s->impl);
    delete s;
}

int pbrs_scene_set_camera(pbrs_scene *s, uint32_t width, uint32_t height, float fov_y_rad, const float eye[3], const float target[3],
                          const float up[3]) {
    NEED(s && eye && target && up, "set_camera: null argument");
    NOT_COMMITTED(s);
    NEED(width > 1 && height > 1, "set_camera: the frame must be at least 2x2 (Camera::new divides by width/2)");
    NEED(std::isfinite(fov_y_rad) && finite3(eye) && finite3(target) && finite3(up), "set_camera: non-finite argument");
    return host_set_camera(s->impl, width, height, fov_y_rad, eye, target, up);
}

int pbrs_scene_add_texture_solid(pbrs_scene *s, const float rgb[3]) {
    NEED(s && rgb, "add_texture_solid: null argument");
    NOT_COMMITTED(s);
    HostTexture t;
    std::memset(&t.rec, 0, sizeof t.rec);
    t.rec.kind = PBRS_TEX_SOLID;
    for (int k = 0; k < 3; ++k) t.rec.value[k] = rgb[k];
    s->impl.textures.push_back(std::move(t));
    return (int)s->impl.textures.size() - 1;
}
int pbrs_scene_add_texture_image_rgb8(pbrs_scene *s, uint32_t width, uint32_t height, const uint8_t *rgb) {
    NEED(s && rgb && width > 0 && height > 0, "add_texture_image: null or empty image");
    NOT_COMMITTED(s);
    s->impl.textures.push_back(make_image(width, height, rgb));
    return (int)s->impl.textures.size() - 1;
}
int pbrs_scene_add_texture_perlin(pbrs_scene *s, float freq, const float rand_vec[256 * 3], const uint32_t perm_x[256],
                                  const uint32_t perm_y[256], const uint32_t perm_z[256]) {
    NEED(s && rand_vec && perm_x && perm_y && perm_z, "add_texture_perlin: null argument");
    NOT_COMMITTED(s);
    HostTexture t;
    std::memset(&t.rec, 0, sizeof t.rec);
    t.rec.kind = PBRS_TEX_PERLIN;
    t.rec.freq = freq;
    t.perlin_vec.assign(rand_vec, rand_vec + 768);
    t.perlin_perm.resize(768);
    for (int i = 0; i < 256; ++i) {
        NEED(perm_x[i] < 256 && perm_y[i] < 256 && perm_z[i] < 256, "add_texture_perlin: permutation entry out of range");
        t.perlin_perm[i] = perm_x[i];
        t.perlin_perm[256 + i] = perm_y[i];
        t.perlin_perm[512 + i] = perm_z[i];
    }
    s->impl.textures.push_back(std::move(t));
    return (int)s->impl.textures.size() - 1;
}

int pbrs_scene_add_material(pbrs_scene *s, const pbrs_material_desc *d) {
    NEED(s && d, "add_material: null argument");
    NOT_COMMITTED(s);
    NEED(d->kind >= PBRS_MTL_LAMBERTIAN && d->kind <= PBRS_MTL_SUBSTRATE, "add_material: unknown kind");
    const int nt = (int)s->impl.textures.size();
    auto tex_ok = [nt](int id, bool required) { return required ? (id >= 0 && id < nt) : (id < nt); };
    switch (d->kind) {
    case PBRS_MTL_LAMBERTIAN: NEED(tex_ok(d->tex_kd, true), "add_material: lambertian needs tex_kd"); break;
    case PBRS_MTL_UBER:
        NEED(tex_ok(d->tex_kd, true) && tex_ok(d->tex_ks, true) && tex_ok(d->tex_kr, false) && tex_ok(d->tex_kt, false),
             "add_material: uber needs tex_kd and tex_ks");
        break;
    case PBRS_MTL_SUBSTRATE: NEED(tex_ok(d->tex_kd, true) && tex_ok(d->tex_ks, true), "add_material: substrate needs tex_kd and tex_ks"); break;
    default: break;
    }
    MaterialRec m;
    std::memset(&m, 0, sizeof m);
    m.kind = d->kind;
    m.tex_kd = d->tex_kd; m.tex_ks = d->tex_ks; m.tex_kr = d->tex_kr; m.tex_kt = d->tex_kt;
    for (int k = 0; k < 3; ++k) { m.a[k] = d->color_a[k]; m.b[k] = d->color_b[k]; }
    for (int k = 0; k < 4; ++k) m.f[k] = d->f[k];
    m.remap = d->remap_roughness ? 1 : 0;
    s->impl.materials.push_back(m);
    return (int)s->impl.materials.size() - 1;
}

int pbrs_scene_add_sphere(pbrs_scene *s, const float center[3], float radius) {
    NEED(s && center, "add_sphere: null argument");
    NOT_COMMITTED(s);
    NEED(finite3(center) && std::isfinite(radius), "add_sphere: non-finite argument");
    SphereRec r;
    for (int k = 0; k < 3; ++k) r.c[k] = center[k];
    r.r = radius;
    s->impl.spheres.push_back(r);
    s->impl.shapes.push_back(HostShape{PBRS_SHAPE_SPHERE, (uint32_t)s->impl.spheres.size() - 1});
    return (int)s->impl.shapes.size() - 1;
}
int pbrs_scene_add_mesh(pbrs_scene *s, const float *P, const float *N, const float *UV, uint32_t nverts, const uint32_t *idx, uint32_t ntris) {
    NEED(s && P && idx, "add_mesh: null argument");
    NOT_COMMITTED(s);
    NEED(nverts > 0 && ntris > 0, "add_mesh: empty mesh");
    return host_add_mesh(s->impl, P, N, UV, nverts, idx, ntris);
}
int pbrs_scene_add_quad(pbrs_scene *s, const float origin[3], const float side_u[3], const float side_v[3]) {
    NEED(s && origin && side_u && side_v, "add_quad: null argument");
    NOT_COMMITTED(s);
    NEED(finite3(origin) && finite3(side_u) && finite3(side_v), "add_quad: non-finite argument");
    return host_add_simple(s->impl, PBRS_SHAPE_QUAD, origin, side_u, side_v);
}
int pbrs_scene_add_cuboid(pbrs_scene *s, const float p0[3], const float p1[3]) {
    NEED(s && p0 && p1, "add_cuboid: null argument");
    NOT_COMMITTED(s);
    NEED(finite3(p0) && finite3(p1), "add_cuboid: non-finite argument");
    float mn[3], mx[3];
    for (int k = 0; k < 3; ++k) {  // float::min_max, math/src/float.rs:197-203
        if (p0[k] < p1[k]) { mn[k] = p0[k]; mx[k] = p1[k]; } else { mn[k] = p1[k]; mx[k] = p0[k]; }
    }
    return host_add_simple(s->impl, PBRS_SHAPE_CUBOID, mn, mx, nullptr);
}
int pbrs_scene_add_disk(pbrs_scene *s, const float center[3], const float normal[3], const float radial[3]) {
    NEED(s && center && normal && radial, "add_disk: null argument");
    NOT_COMMITTED(s);
    NEED(finite3(center), "add_disk: non-finite centre");
    float n[3];
    if (int rc = host_make_disk(center, normal, radial, n)) return rc;
    return host_add_simple(s->impl, PBRS_SHAPE_DISK, center, n, radial);
}
int pbrs_scene_add_triangle(pbrs_scene *s, const float p0[3], const float p1[3], const float p2[3]) {
    NEED(s && p0 && p1 && p2, "add_triangle: null argument");
    NOT_COMMITTED(s);
    NEED(finite3(p0) && finite3(p1) && finite3(p2), "add_triangle: non-finite argument");
    return host_add_simple(s->impl, PBRS_SHAPE_TRIANGLE, p0, p1, p2);
}
int pbrs_scene_add_sphere_blas(pbrs_scene *s, const float *centers_radii, uint32_t n) {
    NEED(s && centers_radii, "add_sphere_blas: null argument");
    NOT_COMMITTED(s);
    NEED(n > 0, "add_sphere_blas: no spheres");
    return host_add_sphere_blas(s->impl, centers_radii, n);
}
int pbrs_scene_add_instance(pbrs_scene *s, int shape_id, int material_id, const float fwd4x4[16], const float inv4x4[16]) {
    NEED(s, "add_instance: null scene");
    NOT_COMMITTED(s);
    NEED(shape_id >= 0 && shape_id < (int)s->impl.shapes.size(), "add_instance: bad shape id");
    NEED(material_id >= 0 && material_id < (int)s->impl.materials.size(), "add_instance: bad material id");
    NEED((fwd4x4 == nullptr) == (inv4x4 == nullptr), "add_instance: give both matrices or neither");
    return host_add_instance(s->impl, shape_id, material_id, fwd4x4, inv4x4);
}

int pbrs_scene_add_point_light(pbrs_scene *s, const float position[3], const float intensity[3]) {
    NEED(s && position && intensity, "add_point_light: null argument");
    NOT_COMMITTED(s);
    DeltaLightRec l;
    std::memset(&l, 0, sizeof l);
    l.kind = PBRS_LIGHT_POINT;
    for (int k = 0; k < 3; ++k) { l.position[k] = position[k]; l.color[k] = intensity[k]; }
    s->impl.delta_lights.push_back(l);
    return 0;
}
int pbrs_scene_add_distant_light(pbrs_scene *s, const float casting_dir[3], const float radiance[3], float world_radius) {
    NEED(s && casting_dir && radiance, "add_distant_light: null argument");
    NOT_COMMITTED(s);
    DeltaLightRec l;
    std::memset(&l, 0, sizeof l);
    l.kind = PBRS_LIGHT_DISTANT;
    for (int k = 0; k < 3; ++k) { l.casting_dir[k] = casting_dir[k]; l.color[k] = radiance[k]; }
    l.world_radius = world_radius;
    s->impl.delta_lights.push_back(l);
    return 0;
}
int pbrs_scene_add_area_light_sphere(pbrs_scene *s, const float center[3], float radius, const float emit[3]) {
    NEED(s && center && emit, "add_area_light_sphere: null argument");
    NOT_COMMITTED(s);
    AreaLightRec l;
    std::memset(&l, 0, sizeof l);
    l.kind = PBRS_AREA_SPHERE;
    for (int k = 0; k < 3; ++k) { l.p0[k] = center[k]; l.emit[k] = emit[k]; }
    l.p1[0] = radius;
    l.area = radius * radius * 4.0f * kPi;  // Sphere::area, light/src/sample_shape.rs:253-255
    s->impl.area_lights.push_back(l);
    return 0;
}
int pbrs_scene_add_area_light_triangle(pbrs_scene *s, const float p0[3], const float p1[3], const float p2[3], const float emit[3]) {
    NEED(s && p0 && p1 && p2 && emit, "add_area_light_triangle: null argument");
    NOT_COMMITTED(s);
    AreaLightRec l;
    std::memset(&l, 0, sizeof l);
    l.kind = PBRS_AREA_TRIANGLE;
    for (int k = 0; k < 3; ++k) { l.p0[k] = p0[k]; l.p1[k] = p1[k]; l.p2[k] = p2[k]; l.emit[k] = emit[k]; }
    // IsolatedTriangle::area, light/src/sample_shape.rs:289-292: |(p0-p1) x (p2-p1)| * 0.5
    float a[3] = {p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]}, b[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    l.area = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) * 0.5f;
    s->impl.area_lights.push_back(l);
    return 0;
}
int pbrs_scene_add_area_light_quad(pbrs_scene *s, const float origin[3], const float side_u[3], const float side_v[3], const float emit[3]) {
    NEED(s && origin && side_u && side_v && emit, "add_area_light_quad: null argument");
    NOT_COMMITTED(s);
    AreaLightRec l;
    std::memset(&l, 0, sizeof l);
    l.kind = PBRS_AREA_QUAD;
    for (int k = 0; k < 3; ++k) { l.p0[k] = origin[k]; l.p1[k] = side_u[k]; l.p2[k] = side_v[k]; l.emit[k] = emit[k]; }
    // ParallelQuad::area, light/src/sample_shape.rs:306-308: |side_u x side_v|
    const float *a = side_u, *b = side_v;
    float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    l.area = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    s->impl.area_lights.push_back(l);
    return 0;
}
int pbrs_scene_add_area_light_disk(pbrs_scene *s, const float center[3], const float normal[3], const float radial[3], const float emit[3]) {
    NEED(s && center && normal && radial && emit, "add_area_light_disk: null argument");
    NOT_COMMITTED(s);
    AreaLightRec l;
    std::memset(&l, 0, sizeof l);
    l.kind = PBRS_AREA_DISK;
    float n[3];
    if (int rc = host_make_disk(center, normal, radial, n)) return rc;
    for (int k = 0; k < 3; ++k) { l.p0[k] = center[k]; l.p1[k] = n[k]; l.p2[k] = radial[k]; l.emit[k] = emit[k]; }
    // Disk::area, light/src/sample_shape.rs:271-273: |radial|^2 * PI
    l.area = (radial[0] * radial[0] + radial[1] * radial[1] + radial[2] * radial[2]) * kPi;
    s->impl.area_lights.push_back(l);
    return 0;
}

int pbrs_scene_set_env_constant(pbrs_scene *s, const float rgb[3]) {
    NEED(s && rgb, "set_env_constant: null argument");
    NOT_COMMITTED(s);
    s->impl.env_kind = PBRS_ENV_KIND_CONSTANT;
    for (int k = 0; k < 3; ++k) s->impl.env_color[k] = rgb[k];
    return 0;
}
int pbrs_scene_set_env_fn(pbrs_scene *s, int kind) {
    NEED(s, "set_env_fn: null scene");
    NOT_COMMITTED(s);
    NEED(kind >= PBRS_ENV_BLUE_SKY && kind <= PBRS_ENV_DUSK, "set_env_fn: unknown kind");
    s->impl.env_kind = PBRS_ENV_KIND_FN;
    s->impl.env_fn = kind;
    return 0;
}
int pbrs_scene_set_env_image(pbrs_scene *s, uint32_t width, uint32_t height, const uint8_t *rgb, const float scale[3]) {
    NEED(s && rgb && scale && width > 0 && height > 0, "set_env_image: null or empty image");
    NOT_COMMITTED(s);
    s->impl.env_kind = PBRS_ENV_KIND_IMAGE;
    s->impl.env_image = make_image(width, height, rgb);
    for (int k = 0; k < 3; ++k) s->impl.env_scale[k] = scale[k];
    return 0;
}

int pbrs_scene_commit(pbrs_scene *s) {
    NEED(s, "commit: null scene");
    NOT_COMMITTED(s);
    if (!s->impl.has_camera) return fail(PBRS_ERR_STATE, "commit: no camera set");
    if (s->impl.instances.empty()) return fail(PBRS_ERR_STATE, "commit: empty instances (tlas/src/bvh.rs:117 asserts)");
    int rc = host_build(s->impl);
    if (rc < 0) return rc;
    rc = device_upload(s->impl);
    if (rc < 0) return rc;
    s->impl.committed = true;
    return 0;
}

int pbrs_scene_get_info(const pbrs_scene *s, pbrs_scene_info *info) {
    NEED(s && info, "get_info: null argument");
    if (!s->impl.committed) return fail(PBRS_ERR_STATE, "get_info before commit");
    *info = s->impl.info;
    return 0;
}

uint32_t pbrs_sampler_u32(uint64_t seed, uint32_t pixel_index, uint32_t sample_index, uint32_t dimension) {
    return sampler_u32(seed, pixel_index, sample_index, dimension);
}

}  // extern "C"
