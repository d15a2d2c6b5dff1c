// The render half of the C ABI of include/pbrs_gpu.h: device buffers for the outputs, the
// wavefront (kernels.cu), the device-to-host copies that make pbrs_render the end-to-end path, and
// the fan-out of one pbrs_render call over several GPUs (pbrs_render_opts::num_gpus).
//
// Film return path.  A rank that renders a tile split owns a lattice of 64x64 tiles; with
// PBRS_FLAG_OWN_TILES_ONLY it copies exactly those tiles (one 2-D copy each: 64 rows of 768 bytes)
// from its device film straight into the caller's row-major host film and touches nothing else.
// Ranks that share one host film -- the worker threads of a num_gpus > 1 call here, or processes
// that map the same shared-memory buffer (bench.py under torchrun) -- assemble the frame with no
// inter-GPU traffic and no second pass over the film.  A sample split has every rank hold a partial
// sum of every pixel: those are added by ONE kernel on the first device that reads the peers' films
// over NVLink (peer access), then the film crosses PCIe once.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "render.h"

using namespace pbrs;

struct pbrs_scene {
    SceneImpl impl;
};

namespace {

int fail(int code, const char *msg) {
    set_error(msg);
    return code;
}
#define NEED(cond, msg)                                      \
    do {                                                     \
        if (!(cond)) return fail(PBRS_ERR_INVALID_ARG, msg); \
    } while (0)

// device scratch that lives for one call
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("out of device memory for the output buffer"); return PBRS_ERR_OOM; }
        return 0;
    }
};
int cuda_fail(cudaError_t e, const char *what) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return PBRS_ERR_CUDA;
}
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } cudaSetDevice(dev); }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

size_t film_bytes_of(const SceneImpl &s) { return sizeof(float) * 3 * (size_t)s.cam.width * s.cam.height; }

// the replica's own device film (current device must be the replica's)
int ensure_film(const SceneImpl &s, Replica &r) {
    const size_t bytes = film_bytes_of(s);
    if (r.film_bytes >= bytes) return 0;
    if (r.film) cudaFree(r.film);
    r.film = nullptr; r.film_bytes = 0;
    cudaError_t e = cudaMalloc(&r.film, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(PBRS_ERR_OOM, "out of device memory for the film"); }
    r.film_bytes = bytes;
    return 0;
}

// Device film -> the caller's host film.  Whole film: one copy.  Own tiles only: one 2-D copy per
// owned tile, all enqueued before the single synchronisation.
int film_to_host(const SceneImpl &s, Replica &r, const pbrs_render_opts &o, float *out_rgb) {
    const uint32_t W = s.cam.width, H = s.cam.height;
    cudaError_t e;
    const bool own_only = (o.flags & PBRS_FLAG_OWN_TILES_ONLY) && o.world_size > 1 && o.split == PBRS_SPLIT_TILES;
    if (!own_only) {
        e = cudaMemcpy(out_rgb, r.film, film_bytes_of(s), cudaMemcpyDeviceToHost);  // synchronises the frame
        return e == cudaSuccess ? 0 : cuda_fail(e, "film copy");
    }
    std::vector<uint32_t> tiles;
    owned_tiles(s, o, tiles);
    const uint32_t tiles_x = (W + 63) / 64;
    const size_t pitch = sizeof(float) * 3 * (size_t)W;
    for (uint32_t t : tiles) {
        const uint32_t x = (t % tiles_x) * 64, y = (t / tiles_x) * 64;
        const uint32_t w = std::min(64u, W - x), h = std::min(64u, H - y);
        const size_t off = 3 * ((size_t)y * W + x);
        e = cudaMemcpy2DAsync(out_rgb + off, pitch, r.film + off, pitch, sizeof(float) * 3 * w, h, cudaMemcpyDeviceToHost, nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "tile copy");
    }
    e = cudaStreamSynchronize(nullptr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "tile copy");
}

// out[i] = (a[i] + p1[i] + ... ) * scale, summed in rank order: the partial films of a sample split.
// The peers' films are read in place over NVLink (peer access) -- no staging copy, no second pass.
struct PeerFilms {
    const float *p[16];
    int n;
};
__global__ void __launch_bounds__(256) k_film_sum(float *film, PeerFilms peers, size_t n4, size_t n, float scale) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 a = reinterpret_cast<const float4 *>(film)[i];
        for (int k = 0; k < peers.n; ++k) {
            const float4 b = reinterpret_cast<const float4 *>(peers.p[k])[i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        reinterpret_cast<float4 *>(film)[i] = a;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t i = n4 * 4; i < n; ++i) {
            float a = film[i];
            for (int k = 0; k < peers.n; ++k) a += peers.p[k][i];
            film[i] = a * scale;
        }
}

void add_stats(pbrs_stats &a, const pbrs_stats &b) {
    a.n_samples += b.n_samples; a.n_rays_extend += b.n_rays_extend; a.n_rays_shadow += b.n_rays_shadow;
    a.n_nodes += b.n_nodes; a.n_tris += b.n_tris; a.n_spheres += b.n_spheres; a.n_instances += b.n_instances;
    for (int k = 0; k < PBRS_NUM_PANIC_KINDS; ++k) a.would_panic[k] += b.would_panic[k];
    a.ms_total = std::max(a.ms_total, b.ms_total);
    a.ms_generate = std::max(a.ms_generate, b.ms_generate); a.ms_extend = std::max(a.ms_extend, b.ms_extend);
    a.ms_shade = std::max(a.ms_shade, b.ms_shade); a.ms_shadow = std::max(a.ms_shadow, b.ms_shadow);
    a.ms_accumulate = std::max(a.ms_accumulate, b.ms_accumulate);
    a.launches += b.launches; a.launches_extend += b.launches_extend; a.launches_shadow += b.launches_shadow;
    for (int k = 0; k < 4; ++k) { a.trav_extend[k] += b.trav_extend[k]; a.trav_shadow[k] += b.trav_shadow[k]; }
}

// One pbrs_render call over the first `n` CUDA devices: src/main.rs:189-235 gets N GPUs from one call.
int render_multi(SceneImpl &impl, const pbrs_render_opts &o, float *out_rgb, pbrs_stats *st) {
    const int n = o.num_gpus;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess) { cudaGetLastError(); n_dev = 0; }
    if (n > n_dev) return fail(PBRS_ERR_NO_DEVICE, "num_gpus exceeds the number of CUDA devices");
    if (n > 16) return fail(PBRS_ERR_INVALID_ARG, "num_gpus: at most 16");
    if (o.world_size > 1) return fail(PBRS_ERR_INVALID_ARG, "num_gpus > 1 cannot be combined with rank / world_size");
    // replica g > 0 lives on the g-th device other than the commit device
    std::vector<int> devs{impl.device};
    for (int d = 0; d < n_dev && (int)devs.size() < n; ++d)
        if (d != impl.device) devs.push_back(d);
    while ((int)impl.extra.size() < n - 1) {
        Replica *r = new Replica();
        int rc = device_upload_replica(impl, *r, devs[impl.extra.size() + 1]);
        if (rc < 0) { delete r; return rc; }
        impl.extra.push_back(r);
    }
    const bool samples = o.split == PBRS_SPLIT_SAMPLES;
    const uint32_t W = impl.cam.width, H = impl.cam.height;
    if (!samples && o.crop_w != 0 && o.crop_h != 0) std::memset(out_rgb, 0, film_bytes_of(impl));  // tiles outside the crop are nobody's

    std::vector<int> rcs(n, 0);
    std::vector<std::string> errs(n);
    std::vector<pbrs_stats> stats(n);
    auto work = [&](int g) {
        Replica &r = g == 0 ? static_cast<Replica &>(impl) : *impl.extra[g - 1];
        DeviceScope scope(r.device);
        pbrs_render_opts og = o;
        og.num_gpus = 0; og.rank = g; og.world_size = n;
        og.flags |= samples ? PBRS_FLAG_RAW_SUM : PBRS_FLAG_OWN_TILES_ONLY;
        int rc = ensure_film(impl, r);
        RenderTargets tg;
        tg.film = r.film;
        if (rc >= 0) rc = render_frame(impl, r, og, tg, nullptr, st ? &stats[g] : nullptr);
        if (rc >= 0 && !samples) rc = film_to_host(impl, r, og, out_rgb);
        if (rc >= 0 && samples && cudaStreamSynchronize(nullptr) != cudaSuccess) { set_error("frame failed on a device"); rc = PBRS_ERR_CUDA; }
        if (rc >= 0) rc = check_last_frame(r);
        rcs[g] = rc;
        if (rc < 0) errs[g] = get_error();
    };
    std::vector<std::thread> threads;
    for (int g = 1; g < n; ++g) threads.emplace_back(work, g);
    work(0);
    for (auto &t : threads) t.join();
    for (int g = 0; g < n; ++g)
        if (rcs[g] < 0) { set_error("device " + std::to_string(devs[g]) + ": " + errs[g]); return rcs[g]; }

    if (samples) {
        DeviceScope scope(impl.device);
        PeerFilms pf;
        pf.n = n - 1;
        std::vector<DevBuf> staged(n);
        for (int g = 1; g < n; ++g) {
            Replica &r = *impl.extra[g - 1];
            int can = 0;
            cudaDeviceCanAccessPeer(&can, impl.device, r.device);
            cudaError_t e = can ? cudaDeviceEnablePeerAccess(r.device, 0) : cudaErrorPeerAccessUnsupported;
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e == cudaSuccess) {
                pf.p[g - 1] = r.film;
            } else {  // no peer mapping: stage the partial film on the first device
                cudaGetLastError();
                int rc = staged[g].alloc(film_bytes_of(impl));
                if (rc < 0) return rc;
                e = cudaMemcpyPeer(staged[g].p, impl.device, r.film, r.device, film_bytes_of(impl));
                if (e != cudaSuccess) return cuda_fail(e, "peer film copy");
                pf.p[g - 1] = (const float *)staged[g].p;
            }
        }
        const size_t nf = (size_t)W * H * 3;
        const float scale = (o.flags & PBRS_FLAG_RAW_SUM) ? 1.0f : 1.0f / (float)(o.msaa * o.msaa);
        k_film_sum<<<1184, 256>>>(impl.film, pf, nf / 4, nf, scale);
        cudaError_t e = cudaMemcpy(out_rgb, impl.film, film_bytes_of(impl), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return cuda_fail(e, "film copy");
    }
    if (st) {
        std::memset(st, 0, sizeof *st);
        for (int g = 0; g < n; ++g) add_stats(*st, stats[g]);
    }
    return 0;
}

}  // namespace

extern "C" {

int pbrs_render_device(const pbrs_scene *s, const pbrs_render_opts *o, float *d_film, void *cuda_stream, pbrs_stats *st) {
    NEED(s && o && d_film, "render_device: null argument");
    NEED(o->num_gpus <= 1, "render_device: one device per call (num_gpus is for pbrs_render)");
    RenderTargets tg;
    tg.film = d_film;
    SceneImpl &impl = const_cast<pbrs_scene *>(s)->impl;
    return render_frame(impl, impl, *o, tg, reinterpret_cast<cudaStream_t>(cuda_stream), st);
}

int pbrs_render(const pbrs_scene *s, const pbrs_render_opts *o, float *out_rgb, pbrs_stats *st) {
    NEED(s && o && out_rgb, "render: null argument");
    if (!s->impl.committed) return fail(PBRS_ERR_STATE, "render before pbrs_scene_commit");
    SceneImpl &impl = const_cast<pbrs_scene *>(s)->impl;
    if (o->num_gpus > 1) return render_multi(impl, *o, out_rgb, st);
    DeviceScope scope(impl.device);
    int rc = ensure_film(impl, impl);
    if (rc < 0) return rc;
    RenderTargets tg;
    tg.film = impl.film;
    tg.host_film = out_rgb;
    rc = render_frame(impl, impl, *o, tg, nullptr, st);
    if (rc < 0) return rc;
    if (tg.host_copied) {  // the bands went home batch by batch: wait for the last one
        cudaError_t e = cudaStreamSynchronize(nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "film copy");
    } else {
        rc = film_to_host(impl, impl, *o, out_rgb);
        if (rc < 0) return rc;
    }
    return check_last_frame(impl);
}

int pbrs_render_ids(const pbrs_scene *s, const pbrs_render_opts *o, uint32_t sample_index, uint32_t *out_inst, uint32_t *out_prim, float *out_t) {
    NEED(s && o, "render_ids: null argument");
    if (!s->impl.committed) return fail(PBRS_ERR_STATE, "render before pbrs_scene_commit");
    SceneImpl &impl = const_cast<pbrs_scene *>(s)->impl;
    DeviceScope scope(impl.device);
    cudaError_t e;
    pbrs_render_opts opts = *o;
    if (opts.msaa == 0) opts.msaa = 1;
    uint32_t cw = opts.crop_w ? opts.crop_w : impl.cam.width, ch = opts.crop_h ? opts.crop_h : impl.cam.height;
    size_t n = (size_t)cw * ch;
    DevBuf bi, bp, bt;
    int rc;
    if ((rc = bi.alloc(n * 4)) < 0 || (rc = bp.alloc(n * 4)) < 0 || (rc = bt.alloc(n * 4)) < 0) return rc;
    RenderTargets tg;
    tg.ids_inst = (uint32_t *)bi.p; tg.ids_prim = (uint32_t *)bp.p; tg.ids_t = (float *)bt.p;
    tg.only_sample = (int32_t)sample_index;
    rc = render_frame(impl, impl, opts, tg, nullptr, nullptr);
    if (rc < 0) return rc;
    if (out_inst && (e = cudaMemcpy(out_inst, bi.p, n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "ids copy");
    if (out_prim && (e = cudaMemcpy(out_prim, bp.p, n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "ids copy");
    if (out_t && (e = cudaMemcpy(out_t, bt.p, n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "ids copy");
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return cuda_fail(e, "render_ids");
    return check_last_frame(impl);
}

int pbrs_render_samples(const pbrs_scene *s, const pbrs_render_opts *o, float *out_rgb_samples, pbrs_stats *st) {
    NEED(s && o && out_rgb_samples, "render_samples: null argument");
    if (!s->impl.committed) return fail(PBRS_ERR_STATE, "render before pbrs_scene_commit");
    SceneImpl &impl = const_cast<pbrs_scene *>(s)->impl;
    DeviceScope scope(impl.device);
    cudaError_t e;
    NEED(o->world_size <= 1, "render_samples: single-rank only");
    uint32_t cw = o->crop_w ? o->crop_w : impl.cam.width, ch = o->crop_h ? o->crop_h : impl.cam.height;
    size_t n = (size_t)cw * ch * o->msaa * o->msaa * 3;
    DevBuf buf;
    int rc = buf.alloc(n * 4);
    if (rc < 0) return rc;
    RenderTargets tg;
    tg.samples = (float *)buf.p;
    rc = render_frame(impl, impl, *o, tg, nullptr, st);
    if (rc < 0) return rc;
    if ((e = cudaMemcpy(out_rgb_samples, buf.p, n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "samples copy");
    return check_last_frame(impl);
}

// After the caller has synchronised the stream of a pbrs_render_device frame: fails loudly if a
// traversal stack overflowed in it (cannot happen with scenes pbrs_scene_commit accepts).
int pbrs_check_last_frame(const pbrs_scene *s) {
    NEED(s, "check_last_frame: null argument");
    SceneImpl &impl = const_cast<pbrs_scene *>(s)->impl;
    DeviceScope scope(impl.device);
    return check_last_frame(impl);
}

// ---- page-locked host memory for the film: the DMA target of the copies above ----
float *pbrs_film_alloc(uint32_t width, uint32_t height) {
    void *p = nullptr;
    const size_t bytes = sizeof(float) * 3 * (size_t)width * height;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        set_error("film_alloc: cannot allocate page-locked host memory");
        return nullptr;
    }
    return static_cast<float *>(p);
}
void pbrs_film_free(float *p) {
    if (p) cudaFreeHost(p);
}
int pbrs_host_register(void *ptr, uint64_t bytes) {
    NEED(ptr && bytes, "host_register: null argument");
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
    return e == cudaSuccess ? 0 : cuda_fail(e, "host_register");
}
int pbrs_host_unregister(void *ptr) {
    NEED(ptr, "host_unregister: null argument");
    cudaError_t e = cudaHostUnregister(ptr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "host_unregister");
}
int pbrs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // extern "C"
