/*
 * pbrs_gpu.h -- C ABI of the B200 (sm_100a) back end for the pbrs path-tracing inner loop.
 *
 * This is the drop-in boundary.  The reference has no FFI/plugin seam of its
 * own: the narrowest one is the integrator fn-pointer `fn(&Scene, Ray, i32) -> Color` picked at
 * src/main.rs:160-163 and called per sample at src/main.rs:205.  The seam cut here is the whole
 * `image_map` computation, src/main.rs:189-235: inputs (Scene, integrator kind, msaa, depth 5),
 * output the row-major `Vec<Color>` that `write_exr` (src/main.rs:42-53, 245) consumes.
 *
 * Because `Scene` holds `Arc<dyn Shape>` / `Arc<dyn Material>` trait objects that cannot be
 * introspected (tlas/src/instance.rs:12-16), the scene is handed over at CONSTRUCTION time:
 * each call below mirrors one constructor the reference's scene builders use
 * (scene/src/loader.rs:164-305 `traverse_world_item`, scene/src/preset.rs).
 *
 * Conventions
 *   - everything is plain C: POD structs, raw pointers, sizes; no C++/torch types.
 *   - the caller owns all input memory; the library copies during the call.
 *   - functions returning `int` return >= 0 on success (an id where documented) and a
 *     negative pbrs_status on failure; pbrs_last_error() has the message (thread-local).
 *   - nothing aborts: reference panics/asserts become either error codes (bad input) or the
 *     `would_panic` counters in pbrs_stats (numerical asserts on the hot path).
 *   - one pbrs_scene may be rendered from one host thread at a time, one frame in flight: the
 *     path workspace, the counters and the cached CUDA graph belong to the scene (per device).
 *     Every call restores the caller's current CUDA device before it returns.
 *   - there is NO CPU fallback: every render entry point fails with PBRS_ERR_NO_DEVICE when no
 *     CUDA device is usable.
 */
#ifndef PBRS_GPU_H
#define PBRS_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBRS_ABI_VERSION 2

typedef enum pbrs_status {
    PBRS_OK = 0,
    PBRS_ERR_INVALID_ARG = -1, /* null pointer, bad id, NaN where the reference asserts !nan */
    PBRS_ERR_STATE = -2,       /* e.g. render before commit, add after commit */
    PBRS_ERR_NO_DEVICE = -3,   /* no usable CUDA device: there is no CPU fallback */
    PBRS_ERR_CUDA = -4,        /* a CUDA runtime call failed; message in pbrs_last_error */
    PBRS_ERR_UNSUPPORTED = -5, /* reference `todo!()`/`unimplemented!()` territory */
    PBRS_ERR_OOM = -6
} pbrs_status;

typedef struct pbrs_scene pbrs_scene; /* opaque: host description + device mirrors */

/* ---- lifetime -------------------------------------------------------------------------- */
pbrs_scene *pbrs_scene_create(void);
void pbrs_scene_destroy(pbrs_scene *);
const char *pbrs_last_error(void);
int pbrs_abi_version(void);

/* ---- camera: geometry/src/camera.rs:19-44 (Camera::new + look_at) ---------------------- */
int pbrs_scene_set_camera(pbrs_scene *, uint32_t width, uint32_t height, float fov_y_rad,
                          const float eye[3], const float target[3], const float up[3]);

/* ---- textures: texture/src/lib.rs ------------------------------------------------------- */
/* Solid (texture/src/lib.rs:19-33). Returns texture id. */
int pbrs_scene_add_texture_solid(pbrs_scene *, const float rgb[3]);
/* Image (texture/src/lib.rs:162-223): 8-bit RGB rows top to bottom, nearest lookup.  */
int pbrs_scene_add_texture_image_rgb8(pbrs_scene *, uint32_t width, uint32_t height,
                                      const uint8_t *rgb);
/* Perlin (texture/src/lib.rs:52-160).  The reference fills its tables from an OS-seeded RNG
 * (:66-96), so the caller supplies them: 256 unit vectors and three permutations of 0..255. */
int pbrs_scene_add_texture_perlin(pbrs_scene *, float freq, const float rand_vec[256 * 3],
                                  const uint32_t perm_x[256], const uint32_t perm_y[256],
                                  const uint32_t perm_z[256]);

/* ---- materials: material/src/lib.rs ----------------------------------------------------- */
typedef enum pbrs_material_kind {
    PBRS_MTL_LAMBERTIAN = 0,    /* :180  tex_kd                                             */
    PBRS_MTL_METAL = 1,         /* :200  color_a = eta, color_b = k, f[0] = fuzziness       */
    PBRS_MTL_GLOSSY = 2,        /* :216  color_a = albedo, f[0] = roughness                 */
    PBRS_MTL_MIRROR = 3,        /* :229  color_a = albedo                                   */
    PBRS_MTL_DIELECTRIC = 4,    /* :265  f[0] = ior, color_a = reflect, color_b = transmit  */
    PBRS_MTL_DIFFUSE_LIGHT = 5, /* :291  color_a = emit                                     */
    PBRS_MTL_PLASTIC = 6,       /* :433  color_a = kd, color_b = ks, f[0] = roughness, remap */
    PBRS_MTL_UBER = 7,          /* :317  tex_kd, tex_ks, tex_kr/-1, tex_kt/-1,
                                         f[0] = rough_u, f[1] = rough_v, f[2] = eta,
                                         f[3] = opacity, remap                              */
    PBRS_MTL_SUBSTRATE = 8      /* :393  tex_kd, tex_ks (degrades to Lambert upstream)      */
} pbrs_material_kind;

typedef struct pbrs_material_desc {
    int32_t kind;                           /* pbrs_material_kind */
    int32_t tex_kd, tex_ks, tex_kr, tex_kt; /* texture ids, -1 = absent */
    float color_a[3];
    float color_b[3];
    float f[4];
    int32_t remap_roughness; /* bool */
} pbrs_material_desc;

int pbrs_scene_add_material(pbrs_scene *, const pbrs_material_desc *); /* -> material id */

/* ---- shapes: shape/src/simple.rs, shape/src/blas.rs ------------------------------------- */
/* Sphere::new (shape/src/simple.rs:16). -> shape id */
int pbrs_scene_add_sphere(pbrs_scene *, const float center[3], float radius);
/* TriangleMesh::from_soa (shape/src/blas.rs:134-159).  P: nverts*3, N: nverts*3 (may be all
 * zero: the geometric normal is then used, blas.rs:171-174), UV: nverts*2, idx: ntris*3.
 * -> shape id */
int pbrs_scene_add_mesh(pbrs_scene *, const float *P, const float *N, const float *UV,
                        uint32_t nverts, const uint32_t *idx, uint32_t ntris);
/* ParallelQuad{origin, side_u, side_v} (shape/src/simple.rs:69-103: new_xy / new_xz / new_yz are
 * spelled out by the caller).  intersect :120-150 (u, v from cross-product norms; the
 * accurate-vs-coarse assert is counted in would_panic[PBRS_PANIC_QUAD]), occludes :151-163
 * (reciprocal t, transcribed).  -> shape id */
int pbrs_scene_add_quad(pbrs_scene *, const float origin[3], const float side_u[3],
                        const float side_v[3]);
/* Cuboid::from_points (shape/src/simple.rs:173-181; intersect :343-411, occludes = the box
 * test :412-415).  -> shape id */
int pbrs_scene_add_cuboid(pbrs_scene *, const float p0[3], const float p1[3]);
/* Disk::new(center, normal, radial) (shape/src/simple.rs:42-52: the normal is normalised;
 * a non-finite radial or |radial . normal| >= 1e-6 is PBRS_ERR_INVALID_ARG where the reference
 * asserts).  intersect :306-326, occludes :328-332 (ignores the ray extent).  -> shape id */
int pbrs_scene_add_disk(pbrs_scene *, const float center[3], const float normal[3],
                        const float radial[3]);
/* IsolatedTriangle::new (shape/src/simple.rs:184-195; intersect :425-427 = intersect_triangle with
 * dpdu = p1 - p0 and (u, v) = the barycentrics, occludes :428-430): the Shape the loader instances
 * once per face of a PLY area light (scene/src/loader.rs:408-433).  -> shape id */
int pbrs_scene_add_triangle(pbrs_scene *, const float p0[3], const float p1[3], const float p2[3]);
/* IsoBlas::<Sphere>::build (shape/src/blas.rs:60-69, traversal :263-275): n spheres as
 * (cx, cy, cz, radius) under one bottom-level BVH.  The primitive id reported for a hit is the
 * sphere's index in this array.  -> shape id */
int pbrs_scene_add_sphere_blas(pbrs_scene *, const float *centers_radii, uint32_t n);

/* ---- instances: tlas/src/instance.rs:12-45 ---------------------------------------------- */
/* fwd/inv are column-major 4x4 (the reference's Mat4 is four column Vec4s, math/src/hcm.rs:477),
 * NULL = identity.  The bottom row must be (0,0,0,1) (geometry/src/transform.rs:277 asserts
 * w == 1).  -> instance id (instances are numbered in insertion order). */
int pbrs_scene_add_instance(pbrs_scene *, int shape_id, int material_id, const float fwd4x4[16],
                            const float inv4x4[16]);

/* ---- lights: light/src/lib.rs ----------------------------------------------------------- */
int pbrs_scene_add_point_light(pbrs_scene *, const float position[3], const float intensity[3]);
/* world_radius <= 0 or non-finite: derived at commit as Scene::from_loader does
 * (scene/src/lib.rs:54-58: half the TLAS bbox diagonal). */
int pbrs_scene_add_distant_light(pbrs_scene *, const float casting_dir[3],
                                 const float radiance[3], float world_radius);
/* DiffuseAreaLight over SamplableShape::Sphere / ::Triangle (light/src/lib.rs:107-172,
 * light/src/sample_shape.rs:38-43), world-space shape.  The emissive INSTANCE that makes the
 * light visible to camera rays is added separately with PBRS_MTL_DIFFUSE_LIGHT, exactly as
 * scene/src/loader.rs:176-195 does. */
int pbrs_scene_add_area_light_sphere(pbrs_scene *, const float center[3], float radius,
                                     const float emit[3]);
int pbrs_scene_add_area_light_triangle(pbrs_scene *, const float p0[3], const float p1[3],
                                       const float p2[3], const float emit[3]);
/* SamplableShape::Quad / ::Disk (light/src/sample_shape.rs:38-43; sample :257-274,296-309;
 * pdf_at is the trait default :28-33 over the shape's own intersect). */
int pbrs_scene_add_area_light_quad(pbrs_scene *, const float origin[3], const float side_u[3],
                                   const float side_v[3], const float emit[3]);
int pbrs_scene_add_area_light_disk(pbrs_scene *, const float center[3], const float normal[3],
                                   const float radial[3], const float emit[3]);

/* ---- environment: scene/src/lib.rs:12-16,96-117; scene/src/preset.rs:25-51 -------------- */
typedef enum pbrs_env_fn {
    PBRS_ENV_BLUE_SKY = 0,
    PBRS_ENV_DARK_ROOM = 1,
    PBRS_ENV_DUSK = 2
} pbrs_env_fn;
int pbrs_scene_set_env_constant(pbrs_scene *, const float rgb[3]);
int pbrs_scene_set_env_fn(pbrs_scene *, int env_fn_kind);
int pbrs_scene_set_env_image(pbrs_scene *, uint32_t width, uint32_t height, const uint8_t *rgb,
                             const float scale[3]);

/* ---- commit: builds TLAS (tlas/src/bvh.rs:116-152) and every BLAS
 *      (shape/src/blas.rs:333-420), flattens them and uploads to the current CUDA device. --- */
int pbrs_scene_commit(pbrs_scene *);

/* ---- render ------------------------------------------------------------------------------ */
typedef enum pbrs_integrator {
    PBRS_INTEGRATOR_DIRECT = 0, /* src/directlighting.rs:14-47 */
    PBRS_INTEGRATOR_PATH = 1    /* src/pathintegrator.rs:9-74  */
} pbrs_integrator;

typedef enum pbrs_split {
    PBRS_SPLIT_TILES = 0,  /* 64x64 tiles, tile t belongs to rank t % world_size */
    PBRS_SPLIT_SAMPLES = 1 /* sample i of every pixel belongs to rank i % world_size */
} pbrs_split;

#define PBRS_FLAG_COUNT_TRAVERSAL 1u /* fill the n_* traversal counters (slower) */
#define PBRS_FLAG_TIME_STAGES 2u     /* fill ms_* with CUDA-event timings per stage */
#define PBRS_FLAG_NO_JITTER 4u       /* jitter (0,0) as the visualizers do, src/main.rs:170 */
#define PBRS_FLAG_RAW_SUM 8u         /* leave the film as the un-normalised sample sum */
#define PBRS_FLAG_NO_GRAPH 16u       /* enqueue a small frame kernel by kernel instead of replaying its CUDA graph */
/* pbrs_render under a tile split: copy ONLY the tiles this rank owns into out_rgb (one 2-D copy per
 * 64x64 tile) and leave every other pixel of the caller's buffer untouched.  Ranks that share one
 * host film -- the device threads of a num_gpus > 1 call, or processes that map the same
 * shared-memory buffer -- assemble the frame without any inter-GPU traffic. */
#define PBRS_FLAG_OWN_TILES_ONLY 32u

typedef struct pbrs_render_opts {
    int32_t integrator;  /* pbrs_integrator */
    uint32_t msaa;       /* spp = msaa*msaa, src/main.rs:197 */
    int32_t max_depth;   /* the reference hard-codes 5, src/main.rs:205 */
    uint64_t seed;       /* counter-based sampler key (DESIGN.md "Sampler") */
    int32_t rank;        /* this process's share of the frame ... */
    int32_t world_size;  /* ... out of world_size (1 = whole frame) */
    int32_t split;       /* pbrs_split */
    uint32_t crop_x, crop_y, crop_w, crop_h; /* crop_w == 0: full frame */
    uint32_t flags;      /* PBRS_FLAG_* */
    uint32_t paths_in_flight; /* 0 = library default (16 Mi paths); values below 64 Ki are raised to 64 Ki */
    /* 0 or 1: the device that was current at pbrs_scene_commit.  N > 1 (pbrs_render only, with
     * rank = 0 and world_size <= 1): this ONE call spreads the frame over N CUDA devices -- the
     * commit device and the N-1 lowest-numbered others; the scene is replicated on first use --
     * splitting it by `split`, and returns the whole film: src/main.rs:189-235 gets N GPUs from
     * one call.  Tiles: every device copies its own tiles straight into out_rgb.  Samples: the
     * partial sums are added by one kernel on the first device over NVLink peer memory. */
    int32_t num_gpus;
} pbrs_render_opts;

#define PBRS_NUM_PANIC_KINDS 16
typedef struct pbrs_stats {
    uint64_t n_samples;     /* integrator invocations */
    uint64_t n_rays_extend; /* calls to tlas.intersect (closest hit) */
    uint64_t n_rays_shadow; /* calls to tlas.occludes (any hit) */
    /* traversal work, SURVEY.md 8(d) definitions (only with PBRS_FLAG_COUNT_TRAVERSAL) */
    uint64_t n_nodes;       /* TLAS+BLAS inner nodes expanded */
    uint64_t n_tris;        /* triangle records tested */
    uint64_t n_spheres;     /* sphere records tested */
    uint64_t n_instances;   /* instance leaves entered */
    uint64_t would_panic[PBRS_NUM_PANIC_KINDS]; /* reference asserts that would have fired */
    double ms_total;        /* device time of the whole render call's GPU work */
    double ms_generate, ms_extend, ms_shade, ms_shadow, ms_accumulate; /* TIME_STAGES */
    uint64_t launches;      /* kernels launched by this call */
    uint64_t launches_extend;
    /* the traversal counters again, split by kernel: {nodes, tris, spheres, instances} */
    uint64_t trav_extend[4]; /* closest-hit walks (extend kernel) */
    uint64_t trav_shadow[4]; /* any-hit walks (shadow kernel)     */
    uint64_t launches_shadow;
} pbrs_stats;

/* would_panic indices */
#define PBRS_PANIC_SPHERE_INSIDE 0 /* Interaction::new normal.wo >= 0 (interaction.rs:24), D1 */
#define PBRS_PANIC_TBN 1           /* with_dpdu asserts (interaction.rs:46-58)              */
#define PBRS_PANIC_HAT 2           /* Vec3::hat on zero / non-finite (hcm.rs:114)            */
#define PBRS_PANIC_BSDF_FRAME 3    /* BSDF::new_frame asserts (src/bsdf.rs:22-29)            */
#define PBRS_PANIC_MESH_UV 4       /* hit_by_uv distance assert (blas.rs:168)                */
#define PBRS_PANIC_EMPTY_BXDFS 5   /* assert!(!bxdfs.is_empty()) (directlighting.rs:82,124)  */
#define PBRS_PANIC_LOG_SAMPLE 6    /* assert!(log_sample.is_finite()) (microfacet.rs:134)    */
#define PBRS_PANIC_FRESNEL 7       /* conductor ratio is_finite asserts (bxdf.rs:383,388)    */
#define PBRS_PANIC_LAMBERT_WO 8    /* assert!(wo.cos_theta() >= 0) (bxdf.rs:561)             */
#define PBRS_PANIC_PERLIN 9        /* noise range asserts (texture/src/lib.rs:134-135)       */
#define PBRS_PANIC_REFRACT 10      /* assert_ge!(cos_theta_i, 0) (hcm.rs:629)                */
#define PBRS_PANIC_MISC 11
#define PBRS_PANIC_STACK 12         /* not a reference assert: a traversal stack overflowed    */
#define PBRS_PANIC_QUAD 13          /* quad accurate_hit vs coarse_hit assert (simple.rs:140-147), Q11 */

/* Renders this rank's share of the frame and returns the film in HOST memory:
 * out_rgb is width*height*3 floats, row-major, row 0 = top (src/main.rs:219-231), already
 * divided by spp (src/main.rs:208).  Pixels not owned by this rank are 0.  Blocking.
 * Host<->device copies are part of the call (this is the `e2e` path of bench.py). */
int pbrs_render(const pbrs_scene *, const pbrs_render_opts *, float *out_rgb,
                pbrs_stats *stats_or_null);

/* Same, but the film stays on the device: d_film is a device pointer to width*height*3 floats
 * on the scene's device; work is enqueued on `cuda_stream` (a cudaStream_t, NULL = default
 * stream) and the call returns without synchronising unless stats are requested. */
int pbrs_render_device(const pbrs_scene *, const pbrs_render_opts *, float *d_film,
                       void *cuda_stream, pbrs_stats *stats_or_null);

/* After the caller has synchronised the stream a pbrs_render_device frame was enqueued on: fails
 * (PBRS_ERR_UNSUPPORTED) if a traversal stack overflowed during that frame.  pbrs_render,
 * pbrs_render_ids and pbrs_render_samples check this themselves.  pbrs_scene_commit rejects every
 * scene whose BVH depths could overflow the stack, so this is defence in depth: the overflow test
 * itself runs in frames rendered with PBRS_FLAG_COUNT_TRAVERSAL (the counting kernels); the plain
 * kernels rely on the commit-time bound. */
int pbrs_check_last_frame(const pbrs_scene *);

/* Page-locked host memory for the film: the DMA target of pbrs_render's device-to-host copies
 * (a pageable out_rgb works too, through the driver's staging copies: about half the speed).
 * pbrs_film_alloc returns width*height*3 floats usable from every device (NULL on failure);
 * pbrs_host_register pins a buffer the caller already owns -- e.g. the Vec<Color> of
 * src/main.rs:219 -- until pbrs_host_unregister (call it before freeing the memory). */
float *pbrs_film_alloc(uint32_t width, uint32_t height);
void pbrs_film_free(float *);
int pbrs_host_register(void *ptr, uint64_t bytes);
int pbrs_host_unregister(void *ptr);
int pbrs_device_count(void); /* usable CUDA devices (0 if none) */

/* Parity side channels (host buffers; any pointer may be NULL).
 * Primary hit of sample `sample_index` of each pixel of the crop: instance id, primitive id
 * (the triangle's index in the caller's idx array, 0 for spheres), ray t.  Miss = 0xFFFFFFFF
 * ids and t = +inf.  Arrays are crop_w*crop_h (full frame if crop_w == 0). */
int pbrs_render_ids(const pbrs_scene *, const pbrs_render_opts *, uint32_t sample_index,
                    uint32_t *out_inst, uint32_t *out_prim, float *out_t);
/* Per-sample radiance of the crop: out is [crop_h][crop_w][spp][3]. */
int pbrs_render_samples(const pbrs_scene *, const pbrs_render_opts *, float *out_rgb_samples,
                        pbrs_stats *stats_or_null);

/* Scene facts after commit. */
typedef struct pbrs_scene_info {
    uint32_t width, height;
    uint32_t n_instances, n_meshes, n_spheres, n_triangles;
    uint32_t n_tlas_nodes, n_blas_nodes; /* inner nodes (64-byte records) */
    uint32_t n_lights;                   /* delta + area + (env ? 1 : 0) */
    uint64_t device_bytes;               /* HBM footprint of the flattened scene */
    float world_min[3], world_max[3];    /* TLAS root box */
} pbrs_scene_info;
int pbrs_scene_get_info(const pbrs_scene *, pbrs_scene_info *);

/* The sampler, exposed so that callers (and tests) can reproduce a draw:
 * 32 random bits for (seed, pixel_index, sample_index, dimension). */
uint32_t pbrs_sampler_u32(uint64_t seed, uint32_t pixel_index, uint32_t sample_index,
                          uint32_t dimension);

#ifdef __cplusplus
}
#endif
#endif /* PBRS_GPU_H */
