// The recorder that sits next to pbrs's own scene builders and mirrors every constructor call into
// the GPU scene.  `Scene` keeps `Arc<dyn Shape>` / `Arc<dyn Material>`
// (tlas/src/instance.rs:12-16), which cannot be introspected after the fact, so the GPU scene is
// recorded WHILE the Rust scene is built: scene/src/loader.rs `SceneLoader` and scene/src/preset.rs
// call the methods below at the sites ffi/loader_rs.patch shows.
//
// Source for a maintainer (this image has no Rust toolchain); every call goes through the generated
// bindings of ffi/pbrs_gpu.rs (= src/gpu_ffi.rs in the reference tree).
use crate::gpu_ffi::*;
use std::ffi::CStr;

pub struct GpuScene {
    raw: *mut pbrs_scene,
}

#[derive(Debug)]
pub struct GpuError {
    pub code: i32,
    pub message: String,
}

fn check(rc: i32) -> Result<i32, GpuError> {
    if rc >= 0 {
        return Ok(rc);
    }
    // never a silent fallback: the library's message travels with the error
    let message = unsafe { CStr::from_ptr(pbrs_last_error()) }.to_string_lossy().into_owned();
    Err(GpuError { code: rc, message })
}

impl GpuScene {
    pub fn new() -> GpuScene {
        GpuScene { raw: unsafe { pbrs_scene_create() } }
    }

    // geometry/src/camera.rs:19-44 (Camera::new + look_at); `fov_y` is the Angle's radians
    pub fn set_camera(&mut self, (w, h): (u32, u32), fov_y_rad: f32, eye: [f32; 3], target: [f32; 3], up: [f32; 3]) -> Result<(), GpuError> {
        check(unsafe { pbrs_scene_set_camera(self.raw, w, h, fov_y_rad, eye.as_ptr(), target.as_ptr(), up.as_ptr()) }).map(|_| ())
    }

    // texture/src/lib.rs: Solid / Image / Perlin -> texture id
    pub fn texture_solid(&mut self, rgb: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_texture_solid(self.raw, rgb.as_ptr()) })
    }
    pub fn texture_image_rgb8(&mut self, w: u32, h: u32, rgb: &[u8]) -> Result<i32, GpuError> {
        assert_eq!(rgb.len(), (w * h * 3) as usize);
        check(unsafe { pbrs_scene_add_texture_image_rgb8(self.raw, w, h, rgb.as_ptr()) })
    }
    pub fn texture_perlin(&mut self, freq: f32, rand_vec: &[f32; 768], px: &[u32; 256], py: &[u32; 256], pz: &[u32; 256]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_texture_perlin(self.raw, freq, rand_vec.as_ptr(), px.as_ptr(), py.as_ptr(), pz.as_ptr()) })
    }

    // material/src/lib.rs: one tagged record per material kind -> material id
    pub fn material(&mut self, desc: &pbrs_material_desc) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_material(self.raw, desc) })
    }

    // shape/src/simple.rs, shape/src/blas.rs -> shape id
    pub fn sphere(&mut self, center: [f32; 3], radius: f32) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_sphere(self.raw, center.as_ptr(), radius) })
    }
    /// TriangleMesh::from_soa: positions / normals / uvs as flat f32 arrays, index triples as u32
    pub fn mesh(&mut self, p: &[f32], n: &[f32], uv: &[f32], idx: &[u32]) -> Result<i32, GpuError> {
        let nverts = (p.len() / 3) as u32;
        check(unsafe { pbrs_scene_add_mesh(self.raw, p.as_ptr(), n.as_ptr(), uv.as_ptr(), nverts, idx.as_ptr(), (idx.len() / 3) as u32) })
    }
    pub fn quad(&mut self, origin: [f32; 3], side_u: [f32; 3], side_v: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_quad(self.raw, origin.as_ptr(), side_u.as_ptr(), side_v.as_ptr()) })
    }
    pub fn cuboid(&mut self, p0: [f32; 3], p1: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_cuboid(self.raw, p0.as_ptr(), p1.as_ptr()) })
    }
    pub fn disk(&mut self, center: [f32; 3], normal: [f32; 3], radial: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_disk(self.raw, center.as_ptr(), normal.as_ptr(), radial.as_ptr()) })
    }
    pub fn triangle(&mut self, p0: [f32; 3], p1: [f32; 3], p2: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_triangle(self.raw, p0.as_ptr(), p1.as_ptr(), p2.as_ptr()) })
    }
    pub fn sphere_blas(&mut self, centers_radii: &[f32]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_sphere_blas(self.raw, centers_radii.as_ptr(), (centers_radii.len() / 4) as u32) })
    }

    // tlas/src/instance.rs:12-45: Instance::new(shape, mtl).with_transform(t).
    // The reference's Mat4 is four column Vec4s (math/src/hcm.rs:477): pass `forward.cols` / `inverse.cols` flattened.
    pub fn instance(&mut self, shape: i32, material: i32, fwd: Option<&[f32; 16]>, inv: Option<&[f32; 16]>) -> Result<i32, GpuError> {
        let f = fwd.map_or(std::ptr::null(), |m| m.as_ptr());
        let i = inv.map_or(std::ptr::null(), |m| m.as_ptr());
        check(unsafe { pbrs_scene_add_instance(self.raw, shape, material, f, i) })
    }

    // light/src/lib.rs
    pub fn point_light(&mut self, position: [f32; 3], intensity: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_point_light(self.raw, position.as_ptr(), intensity.as_ptr()) })
    }
    pub fn distant_light(&mut self, casting_dir: [f32; 3], radiance: [f32; 3], world_radius: f32) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_distant_light(self.raw, casting_dir.as_ptr(), radiance.as_ptr(), world_radius) })
    }
    pub fn area_light_sphere(&mut self, center: [f32; 3], radius: f32, emit: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_area_light_sphere(self.raw, center.as_ptr(), radius, emit.as_ptr()) })
    }
    pub fn area_light_triangle(&mut self, p0: [f32; 3], p1: [f32; 3], p2: [f32; 3], emit: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_area_light_triangle(self.raw, p0.as_ptr(), p1.as_ptr(), p2.as_ptr(), emit.as_ptr()) })
    }
    pub fn area_light_quad(&mut self, origin: [f32; 3], side_u: [f32; 3], side_v: [f32; 3], emit: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_area_light_quad(self.raw, origin.as_ptr(), side_u.as_ptr(), side_v.as_ptr(), emit.as_ptr()) })
    }
    pub fn area_light_disk(&mut self, center: [f32; 3], normal: [f32; 3], radial: [f32; 3], emit: [f32; 3]) -> Result<i32, GpuError> {
        check(unsafe { pbrs_scene_add_area_light_disk(self.raw, center.as_ptr(), normal.as_ptr(), radial.as_ptr(), emit.as_ptr()) })
    }

    // scene/src/lib.rs:12-16,96-117; scene/src/preset.rs:25-51
    pub fn env_constant(&mut self, rgb: [f32; 3]) -> Result<(), GpuError> {
        check(unsafe { pbrs_scene_set_env_constant(self.raw, rgb.as_ptr()) }).map(|_| ())
    }
    pub fn env_fn(&mut self, kind: i32) -> Result<(), GpuError> {
        check(unsafe { pbrs_scene_set_env_fn(self.raw, kind) }).map(|_| ())
    }
    pub fn env_image(&mut self, w: u32, h: u32, rgb: &[u8], scale: [f32; 3]) -> Result<(), GpuError> {
        check(unsafe { pbrs_scene_set_env_image(self.raw, w, h, rgb.as_ptr(), scale.as_ptr()) }).map(|_| ())
    }

    /// Scene::from_loader's last step: builds TLAS / BLAS with the reference's topology and uploads.
    pub fn commit(&mut self) -> Result<(), GpuError> {
        check(unsafe { pbrs_scene_commit(self.raw) }).map(|_| ())
    }

    /// The whole `image_map` computation of src/main.rs:189-235 in ONE call, over `num_gpus` devices
    /// (tiles: every device copies its own tiles straight into `film`; samples: one peer-memory sum).
    /// `film` is width * height * 3 floats, row-major, row 0 = top; pin it once with
    /// `pbrs_host_register` (or take it from `pbrs_film_alloc`) and the copies are plain DMA.
    pub fn render(&self, integrator: i32, msaa: u32, num_gpus: i32, split: i32, film: &mut [f32]) -> Result<(), GpuError> {
        let opts = pbrs_render_opts {
            integrator, msaa, max_depth: 5, seed: 0x5EED,
            rank: 0, world_size: 1, split,
            crop_x: 0, crop_y: 0, crop_w: 0, crop_h: 0,
            flags: 0, paths_in_flight: 0, num_gpus,
        };
        check(unsafe { pbrs_render(self.raw, &opts, film.as_mut_ptr(), std::ptr::null_mut()) }).map(|_| ())
    }
}

impl Drop for GpuScene {
    fn drop(&mut self) {
        unsafe { pbrs_scene_destroy(self.raw) }
    }
}
