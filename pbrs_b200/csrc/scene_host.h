// Host side of the B200 back end: scene description as handed over the C ABI, the BVH builders
// (same topology as the reference's: tlas/src/bvh.rs:116-152, shape/src/blas.rs:333-420) emitting
// the flattened 64-byte node format directly, and the upload to HBM.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pbrs_gpu.h"
#include "records.h"

namespace pbrs {

struct HostBox {
    float mn[3], mx[3];
};

struct HostMesh {
    std::vector<float> P, N, UV;      // nverts*3, nverts*3, nverts*2
    std::vector<uint32_t> idx;        // ntris*3 (caller order)
    std::vector<SphereRec> balls;     // IsoBlas<Sphere>: the primitives are these spheres; P/N/UV/idx stay empty
    // build output
    std::vector<NodeRec> nodes;       // inner nodes, preorder, indices relative to this mesh
    std::vector<uint32_t> order;      // triangle ids in leaf order
    std::vector<uint32_t> leaf_last;  // positions in `order` that end a leaf
    HostBox root_box;
    bool root_is_leaf = false;
    uint32_t depth = 0;               // longest root-to-leaf chain of inner nodes
};

struct HostShape {
    uint32_t kind;   // PBRS_SHAPE_*
    uint32_t index;  // into spheres / simples / meshes
};

struct HostInstance {
    int shape, material;
    float fwd[4][4], inv[4][4];  // [col][row]
    bool identity;
    HostBox box;                 // world-space, geometry/src/transform.rs:287-308
};

struct HostTexture {
    TextureRec rec;
    std::vector<uint32_t> texels;      // RGBA8
    std::vector<float> perlin_vec;     // 768
    std::vector<uint32_t> perlin_perm; // 768
};

struct DeviceArrays;  // opaque (scene_device.cu)
struct Workspace;     // opaque (kernels.cu)

// Everything a scene owns on ONE device: the uploaded records, the path workspace, the device film.
// The scene itself is replica 0 (the device that was current at pbrs_scene_commit); a render with
// num_gpus > 1 adds one replica per further device (api_render.cu).
struct Replica {
    DeviceArrays *dev = nullptr;
    Workspace *workspace = nullptr;
    float *film = nullptr;         // device film of pbrs_render, kept between calls
    size_t film_bytes = 0;
    DeviceScene dscene{};
    int device = -1;
};

struct SceneImpl : Replica {
    bool has_camera = false, committed = false;
    CameraRec cam{};
    std::vector<HostTexture> textures;
    std::vector<MaterialRec> materials;
    std::vector<HostShape> shapes;
    std::vector<SphereRec> spheres;
    std::vector<SimpleRec> simples;   // quads, cuboids, disks (HostShape::kind tells which)
    std::vector<HostMesh> meshes;
    std::vector<HostInstance> instances;
    std::vector<DeltaLightRec> delta_lights;
    std::vector<AreaLightRec> area_lights;
    int env_kind = PBRS_ENV_KIND_CONSTANT, env_fn = 0;
    float env_color[3] = {0, 0, 0}, env_scale[3] = {1, 1, 1};
    HostTexture env_image;

    // flattened
    std::vector<NodeRec> tlas_nodes;
    HostBox tlas_box{};
    bool tlas_root_is_leaf = false;
    uint32_t tlas_depth = 0;
    pbrs_scene_info info{};

    std::vector<Replica *> extra;  // replicas on further devices (num_gpus > 1), in device order
};

// The flattened record arrays (host copies; scene_device.cu uploads them verbatim).
struct FlatScene {
    std::vector<NodeRec> blas_nodes;
    std::vector<TriRec> tris;
    std::vector<MeshRec> meshes;
    std::vector<TriShadeRec> tri_shade;
    std::vector<InstTravRec> trav;
    std::vector<InstShadeRec> shade;
    std::vector<TextureRec> textures;
    std::vector<uint32_t> texels, perlin_perm;
    std::vector<uint32_t> blas_node_parent, blas_leaf_parent, tlas_node_parent, tlas_leaf_parent;
    std::vector<float> perlin_vec;
    TextureRec env_image;
};
void flatten_scene(const SceneImpl &s, FlatScene &f);
void fill_scene_constants(const SceneImpl &s, const FlatScene &f, DeviceScene &ds);

void set_error(const std::string &msg);
const char *get_error();

// scene_host.cpp
int host_set_camera(SceneImpl &s, uint32_t w, uint32_t h, float fov, const float *eye, const float *target, const float *up);
int host_add_mesh(SceneImpl &s, const float *P, const float *N, const float *UV, uint32_t nverts, const uint32_t *idx, uint32_t ntris);
int host_add_sphere_blas(SceneImpl &s, const float *centers_radii, uint32_t n);
int host_add_simple(SceneImpl &s, uint32_t kind, const float a[3], const float b[3], const float c[3]);
int host_make_disk(const float center[3], const float normal[3], const float radial[3], float out_normal[3]);
int host_add_instance(SceneImpl &s, int shape, int mtl, const float *fwd, const float *inv);
int host_build(SceneImpl &s);  // BLAS per mesh, instance boxes, TLAS; fills tlas_nodes etc.
bool host_tri_may_reject(const HostMesh &m, uint32_t t);

// scene_device.cu
int device_upload(SceneImpl &s);                              // to the current device: replica 0
int device_upload_replica(const SceneImpl &s, Replica &r, int device);  // a further copy on `device`
void device_free(SceneImpl &s);                                // every replica

}  // namespace pbrs
