"""Scene description + handle: the Python host-side mirror of the C ABI (include/pbrs_gpu.h).

``SceneDesc`` records constructor calls in the vocabulary of the reference's scene builders
(scene/src/loader.rs:164-305, scene/src/preset.rs): textures, materials, spheres, triangle meshes,
instances, delta/area lights, environment.  ``SceneDesc.realize(api)`` replays them through a
bound C API (the product's ``pbrs_*``; tests also replay into the oracle's ``oracle_*``) and
returns a ``SceneHandle`` whose ``render*`` methods mirror src/main.rs:189-235.
"""
import ctypes as C

import numpy as np

from . import _capi as K


def _f3(v):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))
    return a, a.ctypes.data_as(K.c_float_p)


def _f32s(pair):
    return np.float32(pair[0]), np.float32(pair[1])


class PbrsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pbrs error {code}: {msg}")
        self.code = code


class SceneDesc:
    def __init__(self):
        self.ops = []
        self.n_tex = self.n_mtl = self.n_shape = self.n_inst = 0
        self.width = self.height = 0

    # -- camera (geometry/src/camera.rs:19-44) --
    def set_camera(self, width, height, fov_y_deg, eye, target, up=(0.0, 1.0, 0.0), fov_y_rad=None):
        # f32::to_radians: deg * (PI / 180) evaluated in f32 (fov_y_rad: an Angle already in radians)
        fov = np.float32(fov_y_deg) * (np.float32(np.pi) / np.float32(180.0)) if fov_y_rad is None else np.float32(fov_y_rad)
        self.width, self.height = int(width), int(height)
        self.ops.append(("scene_set_camera", (int(width), int(height), float(fov), eye, target, up)))

    # -- textures (texture/src/lib.rs) --
    def add_texture_solid(self, rgb):
        self.ops.append(("scene_add_texture_solid", (rgb,)))
        self.n_tex += 1
        return self.n_tex - 1

    def add_texture_image(self, rgb8):
        img = np.ascontiguousarray(rgb8, dtype=np.uint8)
        assert img.ndim == 3 and img.shape[2] == 3
        self.ops.append(("scene_add_texture_image_rgb8", (img,)))
        self.n_tex += 1
        return self.n_tex - 1

    def add_texture_perlin(self, freq, rand_vec, perm_x, perm_y, perm_z):
        self.ops.append(("scene_add_texture_perlin", (float(freq),
                         np.ascontiguousarray(rand_vec, dtype=np.float32),
                         np.ascontiguousarray(perm_x, dtype=np.uint32),
                         np.ascontiguousarray(perm_y, dtype=np.uint32),
                         np.ascontiguousarray(perm_z, dtype=np.uint32))))
        self.n_tex += 1
        return self.n_tex - 1

    # -- materials (material/src/lib.rs) --
    def _add_material(self, kind, tex=(-1, -1, -1, -1), a=(0, 0, 0), b=(0, 0, 0), f=(0, 0, 0, 0), remap=False):
        self.ops.append(("scene_add_material", (kind, tuple(tex), tuple(a), tuple(b), tuple(f), bool(remap))))
        self.n_mtl += 1
        return self.n_mtl - 1

    def lambertian(self, albedo_rgb=None, tex=None):
        if tex is None:
            tex = self.add_texture_solid(albedo_rgb)
        return self._add_material(K.MTL_LAMBERTIAN, tex=(tex, -1, -1, -1))

    def metal(self, eta, k, fuzziness):
        return self._add_material(K.MTL_METAL, a=eta, b=k, f=(fuzziness, 0, 0, 0))

    def glossy(self, albedo, roughness):
        return self._add_material(K.MTL_GLOSSY, a=albedo, f=(roughness, 0, 0, 0))

    def mirror(self, albedo):
        return self._add_material(K.MTL_MIRROR, a=albedo)

    def dielectric(self, ior, reflect=(1, 1, 1), transmit=(1, 1, 1)):
        return self._add_material(K.MTL_DIELECTRIC, a=reflect, b=transmit, f=(ior, 0, 0, 0))

    def diffuse_light(self, emit):
        return self._add_material(K.MTL_DIFFUSE_LIGHT, a=emit)

    def plastic(self, kd, ks, roughness, remap_roughness=True):
        return self._add_material(K.MTL_PLASTIC, a=kd, b=ks, f=(roughness, 0, 0, 0), remap=remap_roughness)

    def uber(self, tex_kd, tex_ks, tex_kr=-1, tex_kt=-1, rough_u=0.0, rough_v=0.0, eta=1.5, opacity=1.0,
             remap_roughness=True):
        return self._add_material(K.MTL_UBER, tex=(tex_kd, tex_ks, tex_kr, tex_kt),
                                  f=(rough_u, rough_v, eta, opacity), remap=remap_roughness)

    def substrate(self, tex_kd, tex_ks):
        return self._add_material(K.MTL_SUBSTRATE, tex=(tex_kd, tex_ks, -1, -1))

    # -- shapes --
    def add_sphere(self, center, radius):
        self.ops.append(("scene_add_sphere", (center, float(radius))))
        self.n_shape += 1
        return self.n_shape - 1

    def add_mesh(self, P, idx, N=None, UV=None):
        P = np.ascontiguousarray(P, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        N = np.zeros_like(P) if N is None else np.ascontiguousarray(N, dtype=np.float32).reshape(-1, 3)
        UV = (np.zeros((P.shape[0], 2), np.float32) if UV is None
              else np.ascontiguousarray(UV, dtype=np.float32).reshape(-1, 2))
        assert N.shape == P.shape and UV.shape[0] == P.shape[0]
        self.ops.append(("scene_add_mesh", (P, N, UV, idx)))
        self.n_shape += 1
        return self.n_shape - 1

    def add_quad(self, origin, side_u, side_v):
        """ParallelQuad{origin, side_u, side_v} (shape/src/simple.rs:69-74)."""
        self.ops.append(("scene_add_quad", (origin, side_u, side_v)))
        self.n_shape += 1
        return self.n_shape - 1

    # ParallelQuad::new_xy / new_xz / new_yz, shape/src/simple.rs:76-102 (differences in f32)
    def add_quad_xy(self, x_range, y_range, z):
        (x0, x1), (y0, y1) = _f32s(x_range), _f32s(y_range)
        return self.add_quad((x0, y0, z), (x1 - x0, 0.0, 0.0), (0.0, y1 - y0, 0.0))

    def add_quad_xz(self, x_range, y, z_range):
        (x0, x1), (z0, z1) = _f32s(x_range), _f32s(z_range)
        return self.add_quad((x0, y, z0), (x1 - x0, 0.0, 0.0), (0.0, 0.0, z1 - z0))

    def add_quad_yz(self, x, y_range, z_range):
        (y0, y1), (z0, z1) = _f32s(y_range), _f32s(z_range)
        return self.add_quad((x, y0, z0), (0.0, 0.0, z1 - z0), (0.0, y1 - y0, 0.0))

    def add_cuboid(self, p0, p1):
        self.ops.append(("scene_add_cuboid", (p0, p1)))
        self.n_shape += 1
        return self.n_shape - 1

    def add_disk(self, center, normal, radial):
        self.ops.append(("scene_add_disk", (center, normal, radial)))
        self.n_shape += 1
        return self.n_shape - 1

    def add_triangle(self, p0, p1, p2):
        """IsolatedTriangle::new (shape/src/simple.rs:184-195)."""
        self.ops.append(("scene_add_triangle", (p0, p1, p2)))
        self.n_shape += 1
        return self.n_shape - 1

    def add_sphere_blas(self, centers_radii):
        """IsoBlas::<Sphere>::build: an (n, 4) array of (cx, cy, cz, radius)."""
        cr = np.ascontiguousarray(centers_radii, dtype=np.float32).reshape(-1, 4)
        self.ops.append(("scene_add_sphere_blas", (cr,)))
        self.n_shape += 1
        return self.n_shape - 1

    def add_instance(self, shape, mtl, fwd=None, inv=None):
        """fwd/inv: 4x4 numpy matrices in the usual math (row, col) convention, or None."""
        if fwd is not None:
            fwd = np.asarray(fwd, dtype=np.float32).reshape(4, 4)
            inv = (np.linalg.inv(fwd.astype(np.float64)).astype(np.float32) if inv is None
                   else np.asarray(inv, dtype=np.float32).reshape(4, 4))
        self.ops.append(("scene_add_instance", (int(shape), int(mtl), fwd, inv)))
        self.n_inst += 1
        return self.n_inst - 1

    # -- lights --
    def add_point_light(self, position, intensity):
        self.ops.append(("scene_add_point_light", (position, intensity)))

    def add_distant_light(self, casting_dir, radiance, world_radius=0.0):
        self.ops.append(("scene_add_distant_light", (casting_dir, radiance, float(world_radius))))

    def add_area_light_sphere(self, center, radius, emit):
        self.ops.append(("scene_add_area_light_sphere", (center, float(radius), emit)))

    def add_area_light_triangle(self, p0, p1, p2, emit):
        self.ops.append(("scene_add_area_light_triangle", (p0, p1, p2, emit)))

    def add_area_light_quad(self, origin, side_u, side_v, emit):
        self.ops.append(("scene_add_area_light_quad", (origin, side_u, side_v, emit)))

    def add_area_light_disk(self, center, normal, radial, emit):
        self.ops.append(("scene_add_area_light_disk", (center, normal, radial, emit)))

    # -- environment --
    def set_env_constant(self, rgb):
        self.ops.append(("scene_set_env_constant", (rgb,)))

    def set_env_fn(self, kind):
        self.ops.append(("scene_set_env_fn", (int(kind),)))

    def set_env_image(self, rgb8, scale=(1, 1, 1)):
        self.ops.append(("scene_set_env_image", (np.ascontiguousarray(rgb8, dtype=np.uint8), scale)))

    # -- replay --
    def realize(self, api, commit=True):
        h = SceneHandle(api)
        keep = []
        for name, args in self.ops:
            fn = api[name]
            if name == "scene_set_camera":
                w, hh, fov, eye, tgt, up = args
                e, ep = _f3(eye); t, tp = _f3(tgt); u, up_ = _f3(up)
                rc = fn(h.ptr, w, hh, fov, ep, tp, up_)
            elif name == "scene_add_texture_image_rgb8":
                img = args[0]
                rc = fn(h.ptr, img.shape[1], img.shape[0], img.ctypes.data_as(K.c_u8_p))
            elif name == "scene_add_texture_perlin":
                freq, rv, px, py, pz = args
                rc = fn(h.ptr, freq, rv.ctypes.data_as(K.c_float_p), px.ctypes.data_as(K.c_u32_p),
                        py.ctypes.data_as(K.c_u32_p), pz.ctypes.data_as(K.c_u32_p))
            elif name == "scene_add_material":
                kind, tex, a, b, f, remap = args
                d = K.MaterialDesc()
                d.kind = kind
                d.tex_kd, d.tex_ks, d.tex_kr, d.tex_kt = tex
                d.color_a = (C.c_float * 3)(*a)
                d.color_b = (C.c_float * 3)(*b)
                d.f = (C.c_float * 4)(*f)
                d.remap_roughness = 1 if remap else 0
                rc = fn(h.ptr, C.byref(d))
            elif name == "scene_add_mesh":
                Pm, Nm, UVm, idx = args
                rc = fn(h.ptr, Pm.ctypes.data_as(K.c_float_p), Nm.ctypes.data_as(K.c_float_p),
                        UVm.ctypes.data_as(K.c_float_p), Pm.shape[0], idx.ctypes.data_as(K.c_u32_p), idx.shape[0])
            elif name == "scene_add_sphere_blas":
                cr = args[0]
                rc = fn(h.ptr, cr.ctypes.data_as(K.c_float_p), cr.shape[0])
            elif name == "scene_add_instance":
                shape, mtl, fwd, inv = args
                if fwd is None:
                    rc = fn(h.ptr, shape, mtl, None, None)
                else:  # C ABI is column-major
                    fa = np.ascontiguousarray(fwd.T.reshape(-1)); ia = np.ascontiguousarray(inv.T.reshape(-1))
                    rc = fn(h.ptr, shape, mtl, fa.ctypes.data_as(K.c_float_p), ia.ctypes.data_as(K.c_float_p))
            elif name == "scene_set_env_image":
                img, scale = args
                s, sp = _f3(scale)
                rc = fn(h.ptr, img.shape[1], img.shape[0], img.ctypes.data_as(K.c_u8_p), sp)
            elif name == "scene_set_env_fn":
                rc = fn(h.ptr, args[0])
            else:
                cargs = []
                for a in args:
                    if isinstance(a, float):
                        cargs.append(a)
                    else:
                        arr, ptr = _f3(a)
                        keep.append(arr)
                        cargs.append(ptr)
                rc = fn(h.ptr, *cargs)
            if rc < 0:
                raise PbrsError(rc, f"{name}: {h.last_error()}")
        h.width, h.height = self.width, self.height
        if commit:
            h.commit()
        return h


class SceneHandle:
    """Owns a ``pbrs_scene*`` (or ``oracle_scene*``)."""

    def __init__(self, api):
        self.api = api
        self.ptr = api["scene_create"]()
        if not self.ptr:
            raise PbrsError(K.ERR_OOM, "scene_create returned NULL")
        self.width = self.height = 0

    def last_error(self):
        e = self.api["last_error"]()
        return e.decode() if e else ""

    def _check(self, rc, what):
        if rc < 0:
            raise PbrsError(rc, f"{what}: {self.last_error()}")
        return rc

    def commit(self):
        self._check(self.api["scene_commit"](self.ptr), "scene_commit")

    def close(self):
        if self.ptr:
            self.api["scene_destroy"](self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        i = K.SceneInfo()
        self._check(self.api["scene_get_info"](self.ptr, C.byref(i)), "scene_get_info")
        return i

    @staticmethod
    def make_opts(integrator="path", msaa=1, max_depth=5, seed=0x5EED, rank=0, world_size=1, split="tiles",
                  crop=None, flags=0, paths_in_flight=0, num_gpus=0):
        o = K.RenderOpts()
        o.integrator = K.INTEGRATOR_PATH if integrator == "path" else K.INTEGRATOR_DIRECT
        o.msaa, o.max_depth, o.seed = int(msaa), int(max_depth), int(seed)
        o.rank, o.world_size = int(rank), int(world_size)
        o.split = K.SPLIT_TILES if split == "tiles" else K.SPLIT_SAMPLES
        if crop is not None:
            o.crop_x, o.crop_y, o.crop_w, o.crop_h = [int(c) for c in crop]
        o.flags = int(flags)
        o.paths_in_flight = int(paths_in_flight)
        o.num_gpus = int(num_gpus)
        return o

    def _crop_wh(self, o):
        return (o.crop_w, o.crop_h) if o.crop_w else (self.width, self.height)

    def render(self, want_stats=True, out=None, **kw):
        """Film [H, W, 3] float32 (row 0 = top), stats dict.  src/main.rs:189-235.
        out: an existing [H, W, 3] float32 film to render into (e.g. a page-locked one)."""
        o = self.make_opts(**kw)
        if out is None:
            out = np.zeros((self.height, self.width, 3), np.float32)
        assert out.shape == (self.height, self.width, 3) and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
        st = K.Stats()
        self._check(self.api["render"](self.ptr, C.byref(o), out.ctypes.data_as(K.c_float_p),
                                       C.byref(st) if want_stats else None), "render")
        return out, st.as_dict()

    def render_ids(self, sample_index=0, **kw):
        o = self.make_opts(**kw)
        w, h = self._crop_wh(o)
        inst = np.zeros((h, w), np.uint32); prim = np.zeros((h, w), np.uint32); t = np.zeros((h, w), np.float32)
        self._check(self.api["render_ids"](self.ptr, C.byref(o), int(sample_index), inst.ctypes.data_as(K.c_u32_p),
                                           prim.ctypes.data_as(K.c_u32_p), t.ctypes.data_as(K.c_float_p)), "render_ids")
        return inst, prim, t

    def render_samples(self, **kw):
        """Per-sample radiance [h, w, spp, 3] of the crop, stats dict."""
        o = self.make_opts(**kw)
        w, h = self._crop_wh(o)
        spp = o.msaa * o.msaa
        out = np.zeros((h, w, spp, 3), np.float32)
        st = K.Stats()
        self._check(self.api["render_samples"](self.ptr, C.byref(o), out.ctypes.data_as(K.c_float_p), C.byref(st)),
                    "render_samples")
        return out, st.as_dict()

    def render_device(self, d_film_ptr, stream=None, want_stats=False, **kw):
        """Film stays on the device (d_film_ptr: int device pointer, W*H*3 floats)."""
        o = self.make_opts(**kw)
        st = K.Stats()
        self._check(self.api["render_device"](self.ptr, C.byref(o), C.c_void_p(d_film_ptr),
                                              C.c_void_p(stream) if stream else None,
                                              C.byref(st) if want_stats else None), "render_device")
        return st.as_dict() if want_stats else None
