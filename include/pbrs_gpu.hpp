// pbrs_gpu.hpp -- C++17 host side above the C ABI (pbrs_gpu.h), header only.
//
// The reference is compiled code (Rust) and no Rust toolchain exists in this image, so the host
// side that a pbrs user programs against is mirrored here in C++: the same public names and
// argument meaning as the crates' constructors, sitting on nothing but the `extern "C"` entry
// points of pbrs_gpu.h.  The Rust binding a maintainer would add is in INTEGRATION.md.
//
//   reference item                                              here
//   math::hcm::{Point3, Vec3}, radiometry::color::Color          pbrs::Point3, Vec3, Color
//   math::Angle::{new_deg, new_rad}                              pbrs::Angle
//   geometry::camera::Camera::{new, look_at}  (camera.rs:19-44)  pbrs::Camera
//   geometry::AffineTransform::{identity, translater, rotater,
//       scaler, *}                          (transform.rs:16-194) pbrs::AffineTransform (FP32, same op order)
//   texture::{Solid, Image, Perlin}         (texture/src/lib.rs)  pbrs::tex::*
//   material::{Lambertian, Metal, Glossy, Mirror, Dielectric,
//       DiffuseLight, Plastic, Uber, Substrate}                  pbrs::mtl::*
//   shape::{Sphere, TriangleMesh::from_soa}                      pbrs::shape::*
//   tlas::instance::Instance::{new, with_transform}              pbrs::Instance
//   light::{DeltaLight::{point, distant}, DiffuseAreaLight::new,
//       SamplableShape::{Sphere, Triangle}}                      pbrs::light::*
//   scene::Scene::{new, with_lights, with_fn_env_light,
//       with_const_env_light, with_env_map}  (scene/src/lib.rs)   pbrs::Scene
//   the image_map computation            (src/main.rs:189-235)   pbrs::render(scene, integrator, msaa)
//   write_exr                              (src/main.rs:42-53)    pbrs::write_exr
//
// Error behaviour: where the reference panics on bad input, the C ABI returns an error code and
// this layer throws pbrs::Error (never aborts, never falls back to a CPU path).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pbrs_gpu.h"

namespace pbrs {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error("pbrs_gpu error " + std::to_string(c) + ": " + what), code(c) {}
};
inline int check(int rc, const char *what) {
    if (rc < 0) throw Error(rc, std::string(what) + ": " + pbrs_last_error());
    return rc;
}

struct Vec3 {
    float x = 0, y = 0, z = 0;
    static Vec3 X() { return {1, 0, 0}; }
    static Vec3 Y() { return {0, 1, 0}; }
    static Vec3 Z() { return {0, 0, 1}; }
    const float *data() const { return &x; }
};
using Point3 = Vec3;
inline Vec3 vec3(float x, float y, float z) { return {x, y, z}; }
inline Point3 point3(float x, float y, float z) { return {x, y, z}; }

struct Color {
    float r = 0, g = 0, b = 0;
    static Color gray(float l) { return {l, l, l}; }
    static Color black() { return {0, 0, 0}; }
    static Color white() { return {1, 1, 1}; }
    const float *data() const { return &r; }
};

struct Angle {
    float radian = 0;
    static Angle new_rad(float r) { return {r}; }
    static Angle new_deg(float d) { return {d * (3.14159265358979323846f / 180.0f)}; }  // f32::to_radians
};

// geometry/src/camera.rs:19-44
class Camera {
public:
    Camera(std::pair<uint32_t, uint32_t> resolution, Angle fov_y) : w_(resolution.first), h_(resolution.second), fov_(fov_y) {}
    Camera &look_at(Point3 from, Point3 target, Vec3 up) { eye_ = from; target_ = target; up_ = up; return *this; }
    Camera looking_at(Point3 from, Point3 target, Vec3 up) const { Camera c = *this; c.look_at(from, target, up); return c; }  // camera.rs:46-56
    std::pair<uint32_t, uint32_t> resolution() const { return {w_, h_}; }
private:
    friend class Scene;
    uint32_t w_, h_;
    Angle fov_;
    Point3 eye_{0, 0, 0}, target_{0, 0, 1};
    Vec3 up_{0, 1, 0};
};

// geometry/src/transform.rs:16-194 over math/src/hcm.rs Mat4 (column vectors; FP32, the same
// operation order as the reference, so composed transforms are bit-identical)
class AffineTransform {
public:
    float fwd[4][4], inv[4][4];  // [col][row]
    static AffineTransform identity() {
        AffineTransform t;
        for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) t.fwd[c][r] = t.inv[c][r] = (c == r) ? 1.0f : 0.0f;
        return t;
    }
    static AffineTransform translater(Vec3 v) {
        AffineTransform t = identity();
        t.fwd[3][0] = v.x; t.fwd[3][1] = v.y; t.fwd[3][2] = v.z;
        t.inv[3][0] = -v.x; t.inv[3][1] = -v.y; t.inv[3][2] = -v.z;
        return t;
    }
    static AffineTransform scaler(Vec3 s) {
        AffineTransform t = identity();
        const float v[3] = {s.x, s.y, s.z};
        for (int k = 0; k < 3; ++k) { t.fwd[k][k] = v[k]; t.inv[k][k] = 1.0f / v[k]; }
        return t;
    }
    // math/src/hcm.rs:508-520; the inverse is the transpose (transform.rs:146-152)
    static AffineTransform rotater(Vec3 axis, Angle angle) {
        AffineTransform t = identity();
        const float sin_t = std::sin(angle.radian), cos_t = std::cos(angle.radian);
        const float a[3] = {axis.x, axis.y, axis.z};
        const float aa = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
        const float inv_len = 1.0f / std::sqrt(aa);
        const float ah[3] = {a[0] * inv_len, a[1] * inv_len, a[2] * inv_len};
        for (int i = 0; i < 3; ++i) {
            float base[3] = {0, 0, 0};
            base[i] = 1.0f;
            const float d = base[0] * a[0] + base[1] * a[1] + base[2] * a[2];
            float vc[3], v1[3];
            for (int k = 0; k < 3; ++k) { vc[k] = d * a[k] / aa; v1[k] = base[k] - vc[k]; }
            const float v2[3] = {v1[1] * ah[2] - v1[2] * ah[1], v1[2] * ah[0] - v1[0] * ah[2], v1[0] * ah[1] - v1[1] * ah[0]};
            for (int k = 0; k < 3; ++k) t.fwd[i][k] = vc[k] + v1[k] * cos_t + v2[k] * sin_t;
        }
        for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) t.inv[c][r] = t.fwd[r][c];
        return t;
    }
    AffineTransform inverse() const {
        AffineTransform t;
        std::memcpy(t.fwd, inv, sizeof fwd);
        std::memcpy(t.inv, fwd, sizeof fwd);
        return t;
    }
    // self * rhs -> self.forward * rhs.forward, rhs.inverse * self.inverse (transform.rs:185-194)
    AffineTransform operator*(const AffineTransform &rhs) const {
        AffineTransform t;
        mul(fwd, rhs.fwd, t.fwd);
        mul(rhs.inv, inv, t.inv);
        return t;
    }
    bool is_identity() const {
        const AffineTransform i = identity();
        return std::memcmp(fwd, i.fwd, sizeof fwd) == 0 && std::memcmp(inv, i.inv, sizeof inv) == 0;
    }
private:
    // hcm.rs:539-556: column c of the product = ZERO + (((a.c0*v0 + a.c1*v1) + a.c2*v2) + a.c3*v3)
    static void mul(const float a[4][4], const float b[4][4], float out[4][4]) {
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r)
                out[c][r] = 0.0f + (((a[0][r] * b[c][0] + a[1][r] * b[c][1]) + a[2][r] * b[c][2]) + a[3][r] * b[c][3]);
    }
};
using InstanceTransform = AffineTransform;

// ---- textures -------------------------------------------------------------------------------
struct Texture {
    enum Kind { SolidK, ImageK, PerlinK } kind = SolidK;
    Color value;
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgb8;
    float freq = 1.0f;
    std::vector<float> rand_vec;
    std::vector<uint32_t> perm_x, perm_y, perm_z;
};
using TextureRef = std::shared_ptr<const Texture>;
namespace tex {
struct Solid {
    static TextureRef create(Color c) { auto t = std::make_shared<Texture>(); t->kind = Texture::SolidK; t->value = c; return t; }
};
struct Image {
    // rows top to bottom, 8-bit RGB (texture/src/lib.rs:169-208 decodes the PNG into this)
    static TextureRef from_rgb8(uint32_t w, uint32_t h, std::vector<uint8_t> rgb) {
        if (rgb.size() != size_t(w) * h * 3) throw Error(PBRS_ERR_INVALID_ARG, "Image: wrong buffer size");
        auto t = std::make_shared<Texture>(); t->kind = Texture::ImageK; t->width = w; t->height = h; t->rgb8 = std::move(rgb); return t;
    }
};
struct Perlin {
    // the reference fills its tables from an OS-seeded RNG (texture/src/lib.rs:66-96): the caller supplies them
    static TextureRef with_tables(float freq, std::vector<float> rand_vec, std::vector<uint32_t> px, std::vector<uint32_t> py, std::vector<uint32_t> pz) {
        if (rand_vec.size() != 768 || px.size() != 256 || py.size() != 256 || pz.size() != 256) throw Error(PBRS_ERR_INVALID_ARG, "Perlin: table sizes");
        auto t = std::make_shared<Texture>(); t->kind = Texture::PerlinK; t->freq = freq; t->rand_vec = std::move(rand_vec);
        t->perm_x = std::move(px); t->perm_y = std::move(py); t->perm_z = std::move(pz); return t;
    }
    // Perlin::with_freq (texture/src/lib.rs:66-96) with a seeded generator in place of thread_rng:
    // 256 unit vectors, three permutations built by the same swap loop.
    static TextureRef with_freq(float freq, uint64_t seed = 0x5EED) {
        uint64_t state = seed;
        auto next = [&]() { state = state * 6364136223846793005ull + 1442695040888963407ull; return uint32_t(state >> 33); };
        std::vector<float> rv(768);
        for (int i = 0; i < 256; ++i) {
            float v[3], n2;
            do {
                for (float &c : v) c = float(next() & 0xFFFFFF) / 8388608.0f - 1.0f;
                n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            } while (n2 > 1.0f || n2 < 1e-6f);
            float inv = 1.0f / std::sqrt(n2);
            for (int c = 0; c < 3; ++c) rv[3 * i + c] = v[c] * inv;
        }
        auto perm = [&]() {
            std::vector<uint32_t> p(256);
            for (uint32_t i = 0; i < 256; ++i) p[i] = i;
            for (uint32_t i = 0; i < 256; ++i) std::swap(p[next() % 256u], p[i]);
            return p;
        };
        auto px = perm(), py = perm(), pz = perm();
        return with_tables(freq, std::move(rv), std::move(px), std::move(py), std::move(pz));
    }
};
}  // namespace tex

// ---- materials (material/src/lib.rs) --------------------------------------------------------
struct Material {
    pbrs_material_desc desc{};
    TextureRef kd, ks, kr, kt;
};
using MaterialRef = std::shared_ptr<const Material>;
namespace mtl {
inline std::shared_ptr<Material> make(int kind) {
    auto m = std::make_shared<Material>();
    m->desc.kind = kind; m->desc.tex_kd = m->desc.tex_ks = m->desc.tex_kr = m->desc.tex_kt = -1;
    return m;
}
inline void put(float dst[3], Color c) { dst[0] = c.r; dst[1] = c.g; dst[2] = c.b; }
struct Lambertian {
    static MaterialRef textured(TextureRef albedo) { auto m = make(PBRS_MTL_LAMBERTIAN); m->kd = std::move(albedo); return m; }
    static MaterialRef solid(Color c) { return textured(tex::Solid::create(c)); }
};
struct Metal {
    static MaterialRef from_ior(Color eta, Color eta_k, float fuzziness) {
        auto m = make(PBRS_MTL_METAL); put(m->desc.color_a, eta); put(m->desc.color_b, eta_k); m->desc.f[0] = fuzziness; return m;
    }
};
struct Glossy {
    static MaterialRef create(Color albedo, float roughness) { auto m = make(PBRS_MTL_GLOSSY); put(m->desc.color_a, albedo); m->desc.f[0] = roughness; return m; }
};
struct Mirror {
    static MaterialRef create(Color albedo) { auto m = make(PBRS_MTL_MIRROR); put(m->desc.color_a, albedo); return m; }
};
struct Dielectric {
    static MaterialRef create(float refract_index, Color reflect = Color::white(), Color transmit = Color::white()) {  // ::new(..).with_colors(..)
        auto m = make(PBRS_MTL_DIELECTRIC); m->desc.f[0] = refract_index; put(m->desc.color_a, reflect); put(m->desc.color_b, transmit); return m;
    }
};
struct DiffuseLight {
    static MaterialRef create(Color emit) { auto m = make(PBRS_MTL_DIFFUSE_LIGHT); put(m->desc.color_a, emit); return m; }
};
struct Plastic {
    static MaterialRef create(Color diffuse, Color specular, float roughness, bool remap_roughness = true) {
        auto m = make(PBRS_MTL_PLASTIC); put(m->desc.color_a, diffuse); put(m->desc.color_b, specular); m->desc.f[0] = roughness;
        m->desc.remap_roughness = remap_roughness; return m;
    }
};
struct Uber {
    static MaterialRef create(TextureRef kd, TextureRef ks, TextureRef kr, TextureRef kt, float rough_u, float rough_v, float eta, float opacity, bool remap_roughness) {
        auto m = make(PBRS_MTL_UBER); m->kd = std::move(kd); m->ks = std::move(ks); m->kr = std::move(kr); m->kt = std::move(kt);
        m->desc.f[0] = rough_u; m->desc.f[1] = rough_v; m->desc.f[2] = eta; m->desc.f[3] = opacity; m->desc.remap_roughness = remap_roughness; return m;
    }
};
struct Substrate {
    static MaterialRef create(TextureRef kd, TextureRef ks, float /*rough*/ = 0.1f, bool /*remap*/ = true) {  // degrades to Lambert upstream (:393-420)
        auto m = make(PBRS_MTL_SUBSTRATE); m->kd = std::move(kd); m->ks = std::move(ks); return m;
    }
};
}  // namespace mtl

// ---- shapes (shape/src/simple.rs, shape/src/blas.rs) ----------------------------------------
struct Shape {
    enum Kind { SphereK, MeshK, QuadK, CuboidK, DiskK, SphereBlasK, TriangleK } kind = SphereK;
    Point3 center;                 // sphere / disk centre, quad origin, cuboid corner 0, triangle p0
    float radius = 1.0f;
    Vec3 a{}, b{};                 // quad: side_u, side_v; cuboid: a = corner 1; disk: normal, radial; triangle: p1, p2
    std::vector<float> P, N, UV;   // mesh; sphere BLAS: P holds (cx, cy, cz, r) per sphere
    std::vector<uint32_t> idx;
};
using ShapeRef = std::shared_ptr<const Shape>;
namespace shape {
struct Sphere {
    static ShapeRef create(Point3 center, float radius) { auto s = std::make_shared<Shape>(); s->center = center; s->radius = radius; return s; }
    static ShapeRef from_raw(float x, float y, float z, float radius) { return create({x, y, z}, radius); }
};
struct TriangleMesh {
    // TriangleMesh::from_soa(positions, normals, uvs, index_triples), shape/src/blas.rs:134-159
    static ShapeRef from_soa(std::vector<float> positions, std::vector<float> normals, std::vector<float> uvs, std::vector<uint32_t> index_triples) {
        auto s = std::make_shared<Shape>();
        s->kind = Shape::MeshK;
        const size_t nv = positions.size() / 3;
        if (positions.size() % 3 || index_triples.size() % 3 || index_triples.empty()) throw Error(PBRS_ERR_INVALID_ARG, "TriangleMesh::from_soa: bad array sizes");
        if (normals.empty()) normals.assign(nv * 3, 0.0f);
        if (uvs.empty()) uvs.assign(nv * 2, 0.0f);
        if (normals.size() != nv * 3 || uvs.size() != nv * 2) throw Error(PBRS_ERR_INVALID_ARG, "TriangleMesh::from_soa: attribute sizes");
        s->P = std::move(positions); s->N = std::move(normals); s->UV = std::move(uvs); s->idx = std::move(index_triples);
        return s;
    }
};
// shape/src/simple.rs:69-103
struct ParallelQuad {
    Point3 origin;
    Vec3 side_u, side_v;
    static ParallelQuad new_xy(std::pair<float, float> x, std::pair<float, float> y, float z) {
        return {{x.first, y.first, z}, {x.second - x.first, 0.0f, 0.0f}, {0.0f, y.second - y.first, 0.0f}};
    }
    static ParallelQuad new_xz(std::pair<float, float> x, float y, std::pair<float, float> z) {
        return {{x.first, y, z.first}, {x.second - x.first, 0.0f, 0.0f}, {0.0f, 0.0f, z.second - z.first}};
    }
    static ParallelQuad new_yz(float x, std::pair<float, float> y, std::pair<float, float> z) {
        return {{x, y.first, z.first}, {0.0f, 0.0f, z.second - z.first}, {0.0f, y.second - y.first, 0.0f}};
    }
    operator ShapeRef() const { auto s = std::make_shared<Shape>(); s->kind = Shape::QuadK; s->center = origin; s->a = side_u; s->b = side_v; return s; }
};
// shape/src/simple.rs:166-182
struct Cuboid {
    static ShapeRef from_points(Point3 p0, Point3 p1) { auto s = std::make_shared<Shape>(); s->kind = Shape::CuboidK; s->center = p0; s->a = p1; return s; }
};
// shape/src/simple.rs:33-66 (new_anyspin's make_coord_system is the library's business: pass the radial)
struct Disk {
    Point3 center;
    Vec3 normal, radial;
    static Disk create(Point3 center, Vec3 normal, Vec3 radial) { return {center, normal, radial}; }
    operator ShapeRef() const { auto s = std::make_shared<Shape>(); s->kind = Shape::DiskK; s->center = center; s->a = normal; s->b = radial; return s; }
};
// shape/src/simple.rs:184-195
struct IsolatedTriangle {
    static ShapeRef create(Point3 p0, Point3 p1, Point3 p2) { auto s = std::make_shared<Shape>(); s->kind = Shape::TriangleK; s->center = p0; s->a = p1; s->b = p2; return s; }
};
// IsoBlas::<Sphere>::build, shape/src/blas.rs:60-69
struct IsoBlas {
    struct Ball { Point3 center; float radius; };
    static ShapeRef build(const std::vector<Ball> &balls) {
        auto s = std::make_shared<Shape>();
        s->kind = Shape::SphereBlasK;
        for (const Ball &b : balls) { s->P.push_back(b.center.x); s->P.push_back(b.center.y); s->P.push_back(b.center.z); s->P.push_back(b.radius); }
        return s;
    }
};
}  // namespace shape

// tlas/src/instance.rs:12-45
struct Instance {
    ShapeRef shape;
    MaterialRef mtl;
    InstanceTransform transform = InstanceTransform::identity();
    bool has_transform = false;
    Instance(ShapeRef s, MaterialRef m) : shape(std::move(s)), mtl(std::move(m)) {}
    Instance with_transform(const InstanceTransform &t) const { Instance i = *this; i.transform = t; i.has_transform = !t.is_identity(); return i; }
};

// ---- lights (light/src/lib.rs) --------------------------------------------------------------
namespace light {
struct DeltaLight {
    bool is_point = true;
    Point3 position;
    Vec3 casting_dir;
    Color color;
    float world_radius = 0.0f;
    static DeltaLight point(Point3 position, Color intensity) { DeltaLight l; l.position = position; l.color = intensity; return l; }
    static DeltaLight distant(float world_radius, Vec3 casting_dir, Color radiance) {
        DeltaLight l; l.is_point = false; l.world_radius = world_radius; l.casting_dir = casting_dir; l.color = radiance; return l;
    }
};
struct SamplableShape {  // light/src/sample_shape.rs:38-43
    enum Kind { SphereK, TriangleK, QuadK, DiskK } kind = SphereK;
    Point3 p0, p1, p2;
    float radius = 0.0f;
    static SamplableShape Sphere(Point3 center, float radius) { SamplableShape s; s.p0 = center; s.radius = radius; return s; }
    static SamplableShape Triangle(Point3 p0, Point3 p1, Point3 p2) { SamplableShape s; s.kind = TriangleK; s.p0 = p0; s.p1 = p1; s.p2 = p2; return s; }
    static SamplableShape Quad(const shape::ParallelQuad &q) { SamplableShape s; s.kind = QuadK; s.p0 = q.origin; s.p1 = q.side_u; s.p2 = q.side_v; return s; }
    static SamplableShape Disk(const shape::Disk &d) { SamplableShape s; s.kind = DiskK; s.p0 = d.center; s.p1 = d.normal; s.p2 = d.radial; return s; }
};
struct DiffuseAreaLight {
    Color emit_radiance;
    SamplableShape shape;
    DiffuseAreaLight(Color emit, SamplableShape s) : emit_radiance(emit), shape(s) {}
};
enum class EnvFn { BlueSky = PBRS_ENV_BLUE_SKY, DarkRoom = PBRS_ENV_DARK_ROOM, Dusk = PBRS_ENV_DUSK };  // scene/src/preset.rs:25-51
}  // namespace light

enum class Integrator { Direct = PBRS_INTEGRATOR_DIRECT, Path = PBRS_INTEGRATOR_PATH };
inline const char *to_str(Integrator i) { return i == Integrator::Direct ? "direct" : "path"; }

// scene/src/lib.rs:19-95.  Owns the pbrs_scene handle once committed.
class Scene {
public:
    Scene(std::vector<Instance> instances, Camera camera) : instances_(std::move(instances)), camera_(camera) {}
    Scene(Scene &&o) noexcept { *this = std::move(o); }
    Scene &operator=(Scene &&o) noexcept {
        std::swap(handle_, o.handle_);
        instances_ = std::move(o.instances_); camera_ = o.camera_; delta_ = std::move(o.delta_); area_ = std::move(o.area_);
        env_kind_ = o.env_kind_; env_color_ = o.env_color_; env_fn_ = o.env_fn_; env_image_ = std::move(o.env_image_); env_scale_ = o.env_scale_;
        return *this;
    }
    Scene(const Scene &) = delete;
    ~Scene() { if (handle_) pbrs_scene_destroy(handle_); }

    Scene with_lights(std::vector<light::DeltaLight> delta, std::vector<light::DiffuseAreaLight> area) && { delta_ = std::move(delta); area_ = std::move(area); return std::move(*this); }
    Scene with_fn_env_light(light::EnvFn f) && { env_kind_ = 1; env_fn_ = f; return std::move(*this); }
    Scene with_const_env_light(Color c) && { env_kind_ = 0; env_color_ = c; return std::move(*this); }
    Scene with_env_map(TextureRef image, Color scale) && { env_kind_ = 2; env_image_ = std::move(image); env_scale_ = scale; return std::move(*this); }
    const Camera &camera() const { return camera_; }

    // Replays the description through the C ABI and commits (BVH build + upload).  Idempotent.
    pbrs_scene *commit() {
        if (handle_) return handle_;
        pbrs_scene *s = pbrs_scene_create();
        if (!s) throw Error(PBRS_ERR_OOM, "pbrs_scene_create");
        try {
            check(pbrs_scene_set_camera(s, camera_.w_, camera_.h_, camera_.fov_.radian, camera_.eye_.data(), camera_.target_.data(), camera_.up_.data()), "set_camera");
            std::map<const Texture *, int> tex_ids;
            std::map<const Material *, int> mtl_ids;
            std::map<const Shape *, int> shape_ids;
            auto tex_id = [&](const TextureRef &t) -> int {
                if (!t) return -1;
                auto it = tex_ids.find(t.get());
                if (it != tex_ids.end()) return it->second;
                int id = -1;
                if (t->kind == Texture::SolidK) id = check(pbrs_scene_add_texture_solid(s, t->value.data()), "add_texture_solid");
                else if (t->kind == Texture::ImageK) id = check(pbrs_scene_add_texture_image_rgb8(s, t->width, t->height, t->rgb8.data()), "add_texture_image");
                else id = check(pbrs_scene_add_texture_perlin(s, t->freq, t->rand_vec.data(), t->perm_x.data(), t->perm_y.data(), t->perm_z.data()), "add_texture_perlin");
                tex_ids[t.get()] = id;
                return id;
            };
            for (const Instance &in : instances_) {
                if (!in.shape || !in.mtl) throw Error(PBRS_ERR_INVALID_ARG, "Instance without shape or material");
                auto mi = mtl_ids.find(in.mtl.get());
                if (mi == mtl_ids.end()) {
                    pbrs_material_desc d = in.mtl->desc;
                    d.tex_kd = tex_id(in.mtl->kd); d.tex_ks = tex_id(in.mtl->ks); d.tex_kr = tex_id(in.mtl->kr); d.tex_kt = tex_id(in.mtl->kt);
                    mi = mtl_ids.emplace(in.mtl.get(), check(pbrs_scene_add_material(s, &d), "add_material")).first;
                }
                auto si = shape_ids.find(in.shape.get());
                if (si == shape_ids.end()) {
                    const Shape &sh = *in.shape;
                    int id = -1;
                    switch (sh.kind) {
                    case Shape::SphereK: id = check(pbrs_scene_add_sphere(s, sh.center.data(), sh.radius), "add_sphere"); break;
                    case Shape::MeshK: id = check(pbrs_scene_add_mesh(s, sh.P.data(), sh.N.data(), sh.UV.data(), uint32_t(sh.P.size() / 3), sh.idx.data(), uint32_t(sh.idx.size() / 3)), "add_mesh"); break;
                    case Shape::QuadK: id = check(pbrs_scene_add_quad(s, sh.center.data(), sh.a.data(), sh.b.data()), "add_quad"); break;
                    case Shape::CuboidK: id = check(pbrs_scene_add_cuboid(s, sh.center.data(), sh.a.data()), "add_cuboid"); break;
                    case Shape::DiskK: id = check(pbrs_scene_add_disk(s, sh.center.data(), sh.a.data(), sh.b.data()), "add_disk"); break;
                    case Shape::TriangleK: id = check(pbrs_scene_add_triangle(s, sh.center.data(), sh.a.data(), sh.b.data()), "add_triangle"); break;
                    case Shape::SphereBlasK: id = check(pbrs_scene_add_sphere_blas(s, sh.P.data(), uint32_t(sh.P.size() / 4)), "add_sphere_blas"); break;
                    }
                    si = shape_ids.emplace(in.shape.get(), id).first;
                }
                if (in.has_transform) check(pbrs_scene_add_instance(s, si->second, mi->second, &in.transform.fwd[0][0], &in.transform.inv[0][0]), "add_instance");
                else check(pbrs_scene_add_instance(s, si->second, mi->second, nullptr, nullptr), "add_instance");
            }
            for (const light::DeltaLight &l : delta_) {
                if (l.is_point) check(pbrs_scene_add_point_light(s, l.position.data(), l.color.data()), "add_point_light");
                else check(pbrs_scene_add_distant_light(s, l.casting_dir.data(), l.color.data(), l.world_radius), "add_distant_light");
            }
            for (const light::DiffuseAreaLight &l : area_) {
                const light::SamplableShape &q = l.shape;
                switch (q.kind) {
                case light::SamplableShape::SphereK: check(pbrs_scene_add_area_light_sphere(s, q.p0.data(), q.radius, l.emit_radiance.data()), "add_area_light_sphere"); break;
                case light::SamplableShape::TriangleK: check(pbrs_scene_add_area_light_triangle(s, q.p0.data(), q.p1.data(), q.p2.data(), l.emit_radiance.data()), "add_area_light_triangle"); break;
                case light::SamplableShape::QuadK: check(pbrs_scene_add_area_light_quad(s, q.p0.data(), q.p1.data(), q.p2.data(), l.emit_radiance.data()), "add_area_light_quad"); break;
                case light::SamplableShape::DiskK: check(pbrs_scene_add_area_light_disk(s, q.p0.data(), q.p1.data(), q.p2.data(), l.emit_radiance.data()), "add_area_light_disk"); break;
                }
            }
            if (env_kind_ == 0) check(pbrs_scene_set_env_constant(s, env_color_.data()), "set_env_constant");
            else if (env_kind_ == 1) check(pbrs_scene_set_env_fn(s, int(env_fn_)), "set_env_fn");
            else check(pbrs_scene_set_env_image(s, env_image_->width, env_image_->height, env_image_->rgb8.data(), env_scale_.data()), "set_env_image");
            check(pbrs_scene_commit(s), "commit");
        } catch (...) {
            pbrs_scene_destroy(s);
            throw;
        }
        handle_ = s;
        return handle_;
    }

private:
    pbrs_scene *handle_ = nullptr;
    std::vector<Instance> instances_;
    Camera camera_{{2, 2}, Angle::new_deg(60)};
    std::vector<light::DeltaLight> delta_;
    std::vector<light::DiffuseAreaLight> area_;
    int env_kind_ = 0;
    Color env_color_ = Color::black();
    light::EnvFn env_fn_ = light::EnvFn::BlueSky;
    TextureRef env_image_;
    Color env_scale_ = Color::white();
};

// The image_map computation of src/main.rs:189-235: every pixel, msaa*msaa stratified samples,
// integrator depth 5, box average.  Row-major, row 0 = top -- what write_exr consumes.
// num_gpus > 1: the one call spreads the frame over that many devices (tiles, or samples with
// split = PBRS_SPLIT_SAMPLES) and still returns the whole film.
inline std::vector<Color> render(Scene &scene, Integrator integrator, uint32_t msaa, pbrs_stats *stats = nullptr, uint64_t seed = 0x5EED,
                                 int num_gpus = 1, int split = PBRS_SPLIT_TILES) {
    pbrs_scene *s = scene.commit();
    auto wh = scene.camera().resolution();
    std::vector<Color> film(size_t(wh.first) * wh.second);
    pbrs_render_opts o{};
    o.integrator = int(integrator); o.msaa = msaa; o.max_depth = 5; o.seed = seed; o.rank = 0; o.world_size = 1;
    o.num_gpus = num_gpus; o.split = split;
    check(pbrs_render(s, &o, &film[0].r, stats), "render");
    return film;
}

// src/main.rs:238-243
inline std::string exr_file_name(const std::string &scene_name, Integrator integrator, uint32_t msaa) {
    return scene_name + "-" + to_str(integrator) + "-" + std::to_string(msaa * msaa) + "spp.exr";
}

// src/main.rs:42-53 write_exr: single-part scan-line OpenEXR, three FLOAT channels, uncompressed.
inline void write_exr(const std::string &file_name, const std::vector<Color> &colors, std::pair<uint32_t, uint32_t> wh) {
    const uint32_t w = wh.first, h = wh.second;
    if (colors.size() != size_t(w) * h) throw Error(PBRS_ERR_INVALID_ARG, "write_exr: film size");
    std::string hdr;
    auto put_i32 = [](std::string &s, int32_t v) { s.append(reinterpret_cast<const char *>(&v), 4); };
    auto attr = [&](const char *name, const char *type, const std::string &payload) {
        hdr.append(name); hdr.push_back('\0'); hdr.append(type); hdr.push_back('\0'); put_i32(hdr, int32_t(payload.size())); hdr.append(payload);
    };
    std::string ch;
    for (const char *n : {"B", "G", "R"}) { ch.append(n); ch.push_back('\0'); put_i32(ch, 2); ch.append(4, '\0'); put_i32(ch, 1); put_i32(ch, 1); }
    ch.push_back('\0');
    std::string box; put_i32(box, 0); put_i32(box, 0); put_i32(box, int32_t(w) - 1); put_i32(box, int32_t(h) - 1);
    auto f32s = [](std::initializer_list<float> v) { std::string s; for (float x : v) s.append(reinterpret_cast<const char *>(&x), 4); return s; };
    attr("channels", "chlist", ch);
    attr("compression", "compression", std::string(1, '\0'));
    attr("dataWindow", "box2i", box);
    attr("displayWindow", "box2i", box);
    attr("lineOrder", "lineOrder", std::string(1, '\0'));
    attr("pixelAspectRatio", "float", f32s({1.0f}));
    attr("screenWindowCenter", "v2f", f32s({0.0f, 0.0f}));
    attr("screenWindowWidth", "float", f32s({1.0f}));
    hdr.push_back('\0');
    FILE *f = std::fopen(file_name.c_str(), "wb");
    if (!f) throw Error(PBRS_ERR_INVALID_ARG, "write_exr: cannot open " + file_name);
    const int32_t magic = 20000630; const uint32_t version = 2;
    std::fwrite(&magic, 4, 1, f); std::fwrite(&version, 4, 1, f); std::fwrite(hdr.data(), 1, hdr.size(), f);
    const uint64_t line_bytes = uint64_t(w) * 12, first = 8 + hdr.size() + uint64_t(h) * 8;
    for (uint32_t y = 0; y < h; ++y) { uint64_t off = first + uint64_t(y) * (8 + line_bytes); std::fwrite(&off, 8, 1, f); }
    std::vector<float> line(size_t(w) * 3);
    for (uint32_t y = 0; y < h; ++y) {
        const int32_t yy = int32_t(y), nb = int32_t(line_bytes);
        std::fwrite(&yy, 4, 1, f); std::fwrite(&nb, 4, 1, f);
        for (uint32_t x = 0; x < w; ++x) { const Color &c = colors[size_t(y) * w + x]; line[x] = c.b; line[w + x] = c.g; line[2 * size_t(w) + x] = c.r; }
        std::fwrite(line.data(), 4, line.size(), f);
    }
    std::fclose(f);
}

}  // namespace pbrs
