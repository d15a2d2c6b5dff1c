"""ctypes mirror of include/pbrs_gpu.h: struct layouts and the function signature table.

The same table binds the product library (prefix ``pbrs_``, pbrs_b200/lib/libpbrs_gpu.so) and,
from tests/bench only, the CPU oracle (prefix ``oracle_``), because the oracle deliberately
exposes the same scene-construction API so that one scene description can be replayed into both.
"""
import ctypes as C

c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)
c_u8_p = C.POINTER(C.c_uint8)

NUM_PANIC_KINDS = 16

# pbrs_material_kind
MTL_LAMBERTIAN, MTL_METAL, MTL_GLOSSY, MTL_MIRROR, MTL_DIELECTRIC = 0, 1, 2, 3, 4
MTL_DIFFUSE_LIGHT, MTL_PLASTIC, MTL_UBER, MTL_SUBSTRATE = 5, 6, 7, 8
ENV_BLUE_SKY, ENV_DARK_ROOM, ENV_DUSK = 0, 1, 2
INTEGRATOR_DIRECT, INTEGRATOR_PATH = 0, 1
SPLIT_TILES, SPLIT_SAMPLES = 0, 1
FLAG_COUNT_TRAVERSAL, FLAG_TIME_STAGES, FLAG_NO_JITTER, FLAG_RAW_SUM, FLAG_NO_GRAPH, FLAG_OWN_TILES_ONLY = 1, 2, 4, 8, 16, 32

ERR_INVALID_ARG, ERR_STATE, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_OOM = -1, -2, -3, -4, -5, -6

PANIC_NAMES = [
    "sphere_inside", "tbn", "hat", "bsdf_frame", "mesh_uv", "empty_bxdfs", "log_sample",
    "fresnel", "lambert_wo", "perlin", "refract", "misc", "stack", "quad", "r14", "r15",
]


class MaterialDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("tex_kd", C.c_int32), ("tex_ks", C.c_int32), ("tex_kr", C.c_int32), ("tex_kt", C.c_int32),
        ("color_a", C.c_float * 3),
        ("color_b", C.c_float * 3),
        ("f", C.c_float * 4),
        ("remap_roughness", C.c_int32),
    ]


class RenderOpts(C.Structure):
    _fields_ = [
        ("integrator", C.c_int32),
        ("msaa", C.c_uint32),
        ("max_depth", C.c_int32),
        ("seed", C.c_uint64),
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
        ("split", C.c_int32),
        ("crop_x", C.c_uint32), ("crop_y", C.c_uint32), ("crop_w", C.c_uint32), ("crop_h", C.c_uint32),
        ("flags", C.c_uint32),
        ("paths_in_flight", C.c_uint32),
        ("num_gpus", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_samples", C.c_uint64),
        ("n_rays_extend", C.c_uint64),
        ("n_rays_shadow", C.c_uint64),
        ("n_nodes", C.c_uint64),
        ("n_tris", C.c_uint64),
        ("n_spheres", C.c_uint64),
        ("n_instances", C.c_uint64),
        ("would_panic", C.c_uint64 * NUM_PANIC_KINDS),
        ("ms_total", C.c_double),
        ("ms_generate", C.c_double), ("ms_extend", C.c_double), ("ms_shade", C.c_double),
        ("ms_shadow", C.c_double), ("ms_accumulate", C.c_double),
        ("launches", C.c_uint64),
        ("launches_extend", C.c_uint64),
        ("trav_extend", C.c_uint64 * 4),
        ("trav_shadow", C.c_uint64 * 4),
        ("launches_shadow", C.c_uint64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("would_panic", "trav_extend", "trav_shadow")}
        d["trav_extend"] = [int(v) for v in self.trav_extend]
        d["trav_shadow"] = [int(v) for v in self.trav_shadow]
        d["would_panic"] = {PANIC_NAMES[i]: int(self.would_panic[i]) for i in range(NUM_PANIC_KINDS)
                            if self.would_panic[i]}
        return d


class SceneInfo(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("n_instances", C.c_uint32), ("n_meshes", C.c_uint32), ("n_spheres", C.c_uint32),
        ("n_triangles", C.c_uint32),
        ("n_tlas_nodes", C.c_uint32), ("n_blas_nodes", C.c_uint32),
        ("n_lights", C.c_uint32),
        ("device_bytes", C.c_uint64),
        ("world_min", C.c_float * 3), ("world_max", C.c_float * 3),
    ]


P = C.c_void_p  # opaque scene handle
f3 = c_float_p

# name (without prefix) -> (restype, argtypes)
SCENE_API = {
    "scene_create": (P, []),
    "scene_destroy": (None, [P]),
    "last_error": (C.c_char_p, []),
    "scene_set_camera": (C.c_int, [P, C.c_uint32, C.c_uint32, C.c_float, f3, f3, f3]),
    "scene_add_texture_solid": (C.c_int, [P, f3]),
    "scene_add_texture_image_rgb8": (C.c_int, [P, C.c_uint32, C.c_uint32, c_u8_p]),
    "scene_add_texture_perlin": (C.c_int, [P, C.c_float, f3, c_u32_p, c_u32_p, c_u32_p]),
    "scene_add_material": (C.c_int, [P, C.POINTER(MaterialDesc)]),
    "scene_add_sphere": (C.c_int, [P, f3, C.c_float]),
    "scene_add_mesh": (C.c_int, [P, f3, f3, f3, C.c_uint32, c_u32_p, C.c_uint32]),
    "scene_add_quad": (C.c_int, [P, f3, f3, f3]),
    "scene_add_cuboid": (C.c_int, [P, f3, f3]),
    "scene_add_disk": (C.c_int, [P, f3, f3, f3]),
    "scene_add_triangle": (C.c_int, [P, f3, f3, f3]),
    "scene_add_sphere_blas": (C.c_int, [P, f3, C.c_uint32]),
    "scene_add_instance": (C.c_int, [P, C.c_int, C.c_int, f3, f3]),
    "scene_add_point_light": (C.c_int, [P, f3, f3]),
    "scene_add_distant_light": (C.c_int, [P, f3, f3, C.c_float]),
    "scene_add_area_light_sphere": (C.c_int, [P, f3, C.c_float, f3]),
    "scene_add_area_light_triangle": (C.c_int, [P, f3, f3, f3, f3]),
    "scene_add_area_light_quad": (C.c_int, [P, f3, f3, f3, f3]),
    "scene_add_area_light_disk": (C.c_int, [P, f3, f3, f3, f3]),
    "scene_set_env_constant": (C.c_int, [P, f3]),
    "scene_set_env_fn": (C.c_int, [P, C.c_int]),
    "scene_set_env_image": (C.c_int, [P, C.c_uint32, C.c_uint32, c_u8_p, f3]),
    "scene_commit": (C.c_int, [P]),
    "render": (C.c_int, [P, C.POINTER(RenderOpts), f3, C.POINTER(Stats)]),
    "render_ids": (C.c_int, [P, C.POINTER(RenderOpts), C.c_uint32, c_u32_p, c_u32_p, f3]),
    "render_samples": (C.c_int, [P, C.POINTER(RenderOpts), f3, C.POINTER(Stats)]),
    "scene_get_info": (C.c_int, [P, C.POINTER(SceneInfo)]),
    "sampler_u32": (C.c_uint32, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]),
}

# entry points only the product has
PRODUCT_ONLY_API = {
    "render_device": (C.c_int, [P, C.POINTER(RenderOpts), C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "abi_version": (C.c_int, []),
    "check_last_frame": (C.c_int, [P]),
    "film_alloc": (C.c_void_p, [C.c_uint32, C.c_uint32]),
    "film_free": (None, [C.c_void_p]),
    "host_register": (C.c_int, [C.c_void_p, C.c_uint64]),
    "host_unregister": (C.c_int, [C.c_void_p]),
    "device_count": (C.c_int, []),
}

# every symbol include/pbrs_gpu.h declares (checked by tests/test_abi.py)
HEADER_SYMBOLS = ["pbrs_" + n for n in list(SCENE_API) + list(PRODUCT_ONLY_API)]


def bind(lib, prefix, table):
    """Attach restype/argtypes; returns {short name: function}. Raises if a symbol is missing."""
    out = {}
    for name, (restype, argtypes) in table.items():
        fn = getattr(lib, prefix + name)
        fn.restype = restype
        fn.argtypes = argtypes
        out[name] = fn
    return out
