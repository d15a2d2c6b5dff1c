// Flattened scene records as they sit in HBM (DESIGN.md "Data layout").  Plain POD, shared by
// the host flattener (scene_host.cpp) and the device kernels (kernels.cu).
//
// Sizes are the ones SURVEY.md 8(d) prices a ray with:
//   inner node   64 B  (both children's boxes + links: one fetch per expansion)
//   triangle     48 B  (3 x float4 positions, ids in the pad lanes) + 16 B: the unit geometric normal
//                      every test of the triangle starts from (precomputed at commit, see TriRec)
//   sphere       16 B  (centre + radius)
//   instance    128 B  (64 B traversal half: inverse 3x4 + shape link;
//                       64 B shading half: forward 3x4 + material link)
#pragma once
#include <stdint.h>

namespace pbrs {

// Child link encoding in NodeRec::meta:
//   bits 0..1   split axis (BLAS only; tlas/src/bvh.rs keeps no axis)
//   bit  2      left child is a leaf      bit 3  right child is a leaf
//   bits 4..17  left leaf primitive count  bits 18..31 right leaf primitive count (informative)
// For a leaf child, child[k] = PBRS_LEAF_BIT | first primitive (BLAS: index into the mesh's
// triangle records, relative to the mesh's first record; the leaf runs up to and including the
// first record carrying PBRS_TRI_LAST_IN_LEAF.  TLAS: instance id).  For an inner child,
// child[k] = node index (BLAS: relative to the mesh's first node).  The leaf bit is baked into the
// link at commit so that the walk's `next` register is the link as loaded.
struct NodeRec {
    float lmin[3], lmax[3];
    float rmin[3], rmax[3];
    uint32_t child[2];
    uint32_t meta;
    uint32_t pad;
};
static_assert(sizeof(NodeRec) == 64, "NodeRec must be 64 bytes");
// Entries of the traversal stack (device_walk.cuh PBRS_WALK_STACK); commit checks the BVH depths against it.
#define PBRS_WALK_STACK_ENTRIES 128
#define PBRS_NODE_LEFT_LEAF 4u
#define PBRS_NODE_RIGHT_LEAF 8u
#define PBRS_MAX_LEAF_PRIMS 16383u
// A BLAS leaf link also carries the run length when it is short: bits 28..30 of child[k] hold the
// number of primitives (1..6; 0 = longer, walk the run by its PBRS_TRI_LAST_IN_LEAF flag), bits
// 0..27 the first primitive.  The any-hit kernel uses it to spread the triangle tests of all the
// lanes that stand at a leaf over the whole warp.
#define PBRS_MANY_INSTANCES 1024u
#define PBRS_LEAF_FIRST_MASK 0x0FFFFFFFu
#define PBRS_LEAF_COUNT_SHIFT 28
#define PBRS_LEAF_COUNT_MAX 6u

// p0/p1/p2 already carry the reference's (i, k, j) vertex swap (shape/src/blas.rs:162-163):
// p0 = pos[idx.0], p1 = pos[idx.2], p2 = pos[idx.1].
// n = hat((p0 - p1) x (p2 - p1)): the unit normal shape/src/simple.rs:441,481 recompute for every
// test.  It depends on the triangle alone, so commit evaluates it once with the same IEEE FP32
// operations (scene_host.cpp, -ffp-contract=off) and the tests start from it: bit-identical, and
// a cross product, a square root and a division fewer per test.  A triangle whose normal cannot
// be normalised (try_hat fails: the reference returns None) carries PBRS_TRI_DEGENERATE.
struct TriRec {
    float p0[3];
    uint32_t orig;   // triangle index in the caller's idx array (the primitive id)
    float p1[3];
    uint32_t flags;  // PBRS_TRI_*
    float p2[3];
    uint32_t pad;
    float n[3];
    uint32_t pad2;
};
static_assert(sizeof(TriRec) == 64, "TriRec must be 64 bytes");
// The hit can be rejected by TriangleMesh::intersect_triangle's tangent check
// (shape/src/blas.rs:193-201); the traversal must evaluate the shading interpolation for it.
#define PBRS_TRI_CHECK_SHADING 1u
#define PBRS_TRI_LAST_IN_LEAF 2u
// The record is a sphere of an IsoBlas<Sphere> (shape/src/blas.rs:36-70): p0 = centre, p1[0] = radius.
#define PBRS_TRI_SPHERE 4u
#define PBRS_TRI_DEGENERATE 8u   // try_hat of the geometric normal fails: every test misses
// A vertex coordinate beyond 1e18 in magnitude: the walk's shortcut for the "interpolated hit
// position is NaN" rejection (simple.rs:467-469) is not provably safe, the full test runs instead.
#define PBRS_TRI_HUGE 16u

// Shading attributes of one triangle, gathered per TRIANGLE (same index as its TriRec) instead of
// per vertex: one dependent fetch after the hit record instead of three (index triple, then the
// vertex arrays).  Vertex order as in TriRec: 0 = idx.0, 1 = idx.2, 2 = idx.1.
struct TriShadeRec {
    float n0[3], n1[3], n2[3];
    float uv0[2], uv1[2], uv2[2];
    float pad;
};
static_assert(sizeof(TriShadeRec) == 64, "TriShadeRec must be 64 bytes");

struct SphereRec {
    float c[3];
    float r;
};
static_assert(sizeof(SphereRec) == 16, "SphereRec must be 16 bytes");

// ParallelQuad / Cuboid / Disk (shape/src/simple.rs:33-182): three float3 each.
//   quad:   a = origin, b = side_u, c = side_v
//   cuboid: a = min,    b = max
//   disk:   a = centre, b = normal (unit), c = radial
//   isolated triangle: a = p0, b = p1, c = p2
struct SimpleRec {
    float a[3]; uint32_t pad0;
    float b[3]; uint32_t pad1;
    float c[3]; uint32_t pad2;
};
static_assert(sizeof(SimpleRec) == 48, "SimpleRec must be 48 bytes");

#define PBRS_SHAPE_SPHERE 0u
#define PBRS_SHAPE_MESH 1u   // a triangle mesh or a sphere BLAS (TriRec flags tell which)
#define PBRS_SHAPE_QUAD 2u
#define PBRS_SHAPE_CUBOID 3u
#define PBRS_SHAPE_DISK 4u
#define PBRS_SHAPE_TRIANGLE 5u  // IsolatedTriangle: a = p0, b = p1, c = p2

// Row r of a 3x4 matrix = (c0[r], c1[r], c2[r], c3[r]) of the reference's column Mat4.
struct InstTravRec {
    float inv[3][4];
    uint32_t shape_kind;
    uint32_t shape_index;  // sphere / simple-shape record index, or mesh header index
    uint32_t identity;     // fwd == inv == I (add_instance got NULL matrices)
    uint32_t pad;
};
static_assert(sizeof(InstTravRec) == 64, "InstTravRec must be 64 bytes");
struct InstShadeRec {
    float fwd[3][4];
    uint32_t material;
    uint32_t cls;  // PBRS_CLS_* of the material: which shade queue a hit on this instance joins
    uint32_t uses_uv;  // the material reads an image texture, i.e. the hit's (u, v) are consumed
    uint32_t pad;
};
static_assert(sizeof(InstShadeRec) == 64, "InstShadeRec must be 64 bytes");

// Per mesh: root box (tested with the incoming t_max, shape/src/blas.rs:428) and array bases.
struct MeshRec {
    float bmin[3], bmax[3];
    uint32_t node_base;   // first NodeRec of this mesh in the BLAS node array
    uint32_t tri_base;    // first TriRec
    uint32_t n_tris;
    uint32_t root_is_leaf;
    uint32_t pad[6];
};
static_assert(sizeof(MeshRec) == 64, "MeshRec must be 64 bytes");

// Material classes: one shade queue and one specialised shade kernel each (DESIGN.md "Wavefront").
#define PBRS_CLS_ANY (-1)        // not specialised (host-sim / generic code)
#define PBRS_CLS_MISS 0          // the ray left the scene
#define PBRS_CLS_EMISSIVE 1      // DiffuseLight: no lobes
#define PBRS_CLS_LAMBERT 2       // Lambertian, Substrate: one Lambert lobe
#define PBRS_CLS_MICROFACET 3    // Metal, Glossy: one Torrance-Sparrow lobe
#define PBRS_CLS_SPECULAR 4      // Mirror, Dielectric: one specular lobe
#define PBRS_CLS_MULTI 5         // Plastic, Uber: several lobes
#define PBRS_NUM_CLS 6

#define PBRS_TEX_SOLID 0
#define PBRS_TEX_IMAGE 1
#define PBRS_TEX_PERLIN 2
struct TextureRec {
    int32_t kind;
    float value[3];
    uint32_t width, height;
    uint32_t texel_base;   // into the RGBA8 texel array
    float freq;
    uint32_t perlin_base;  // into the perlin table arrays (x256)
    uint32_t pad[3];
};

struct MaterialRec {
    int32_t kind;  // pbrs_material_kind
    int32_t tex_kd, tex_ks, tex_kr, tex_kt;
    float a[3], b[3];
    float f[4];
    int32_t remap;
};

#define PBRS_LIGHT_POINT 0
#define PBRS_LIGHT_DISTANT 1
struct DeltaLightRec {
    int32_t kind;
    float position[3];
    float color[3];  // point: intensity; distant: radiance
    float world_radius;
    float casting_dir[3];
    uint32_t pad;
};
#define PBRS_AREA_SPHERE 0
#define PBRS_AREA_TRIANGLE 1
#define PBRS_AREA_QUAD 2      // p0 = origin, p1 = side_u, p2 = side_v
#define PBRS_AREA_DISK 3      // p0 = centre, p1 = normal (unit), p2 = radial
struct AreaLightRec {
    int32_t kind;
    float p0[3];  // sphere: centre
    float p1[3];  // sphere: (radius, -, -)
    float p2[3];
    float emit[3];
    float area;
    uint32_t pad[2];
};

#define PBRS_ENV_KIND_CONSTANT 0
#define PBRS_ENV_KIND_FN 1
#define PBRS_ENV_KIND_IMAGE 2

// Camera constants (geometry/src/camera.rs:65-77): orientation * {a, b, c} precomputed on the host.
struct CameraRec {
    float center[3];
    float a[3], b[3], c[3];
    uint32_t width, height;
};

// Everything a kernel needs, passed by value.
struct DeviceScene {
    const NodeRec *tlas_nodes;
    const NodeRec *blas_nodes;
    const TriRec *tris;
    const SphereRec *spheres;
    const SimpleRec *simples;   // quads, cuboids and disks
    const InstTravRec *inst_trav;
    const InstShadeRec *inst_shade;
    const MeshRec *meshes;
    const TriShadeRec *tri_shade;  // per triangle, same index as `tris`
    // Parent links, read only by the rare exact re-test of a stacked child (device_walk.cuh):
    // node (| bit 31 = the child is the right one) whose record holds the child's box.
    const uint32_t *blas_node_parent;  // per BLAS inner node (same index as blas_nodes)
    const uint32_t *blas_leaf_parent;  // per triangle record: valid at the first record of a leaf
    const uint32_t *tlas_node_parent;  // per TLAS inner node
    const uint32_t *tlas_leaf_parent;  // per instance
    const MaterialRec *materials;
    const TextureRec *textures;
    const uint32_t *texels;     // RGBA8
    const float *perlin_vec;    // 3 x 256 per perlin texture
    const uint32_t *perlin_perm;  // 3 x 256 per perlin texture (x, y, z)
    const DeltaLightRec *delta_lights;
    const AreaLightRec *area_lights;
    uint32_t n_delta, n_area, has_env;
    int32_t env_kind, env_fn;
    float env_color[3];
    float env_scale[3];
    TextureRec env_image;
    CameraRec cam;
    float tlas_min[3], tlas_max[3];
    uint32_t tlas_root_is_leaf;  // a single instance
    uint32_t n_instances;
    uint32_t leaf_vote; // lanes waiting at a leaf that end the node phase of the closest-hit walk (kernels.cu)
    uint32_t has_mesh; // any BLAS at all (a scene of spheres skips the cooperative leaf phase)
    uint32_t shade_split;  // the path integrator shades the heavy classes with k_surface + k_scatter (scenes of big meshes)
    uint32_t coop_closest; // closest-hit walks test their leaf runs cooperatively (scenes with >= 1024 triangles)
    uint32_t has_ext;  // any quad / cuboid / disk instance or sphere BLAS: selects the EXT traversal kernels
    uint32_t cls_mask; // bit c: some instance's material is of shade class c (the miss class always is): absent classes' kernels are not launched
};

}  // namespace pbrs
