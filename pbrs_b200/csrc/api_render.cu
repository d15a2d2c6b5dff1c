// The render half of the C ABI of include/pbrs_gpu.h: device buffers for the outputs, the
// wavefront (kernels.cu), and the device-to-host copies that make pbrs_render the end-to-end path.
#include <cuda_runtime.h>

#include <cstring>
#include <string>

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

}  // namespace

extern "C" {

int pbrs_render_device(const pbrs_scene *s, const pbrs_render_opts *o, float *d_film, void *cuda_stream, pbrs_stats *st) {
    NEED(s && o && d_film, "render_device: null argument");
    RenderTargets tg;
    tg.film = d_film;
    return render_frame(const_cast<pbrs_scene *>(s)->impl, *o, tg, reinterpret_cast<cudaStream_t>(cuda_stream), st);
}

int pbrs_render(const pbrs_scene *s, const pbrs_render_opts *o, float *out_rgb, pbrs_stats *st) {
    NEED(s && o && out_rgb, "render: null argument");
    if (!s->impl.committed) return fail(PBRS_ERR_STATE, "render before pbrs_scene_commit");
    SceneImpl &impl = const_cast<pbrs_scene *>(s)->impl;
    cudaError_t e = cudaSetDevice(impl.device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    size_t bytes = sizeof(float) * 3 * (size_t)impl.cam.width * impl.cam.height;
    if (impl.film_bytes < bytes) {
        if (impl.film) cudaFree(impl.film);
        impl.film = nullptr; impl.film_bytes = 0;
        e = cudaMalloc(&impl.film, bytes);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(PBRS_ERR_OOM, "out of device memory for the film"); }
        impl.film_bytes = bytes;
    }
    RenderTargets tg;
    tg.film = impl.film;
    int rc = render_frame(impl, *o, tg, nullptr, st);
    if (rc < 0) return rc;
    e = cudaMemcpy(out_rgb, impl.film, bytes, cudaMemcpyDeviceToHost);  // synchronises the frame
    if (e != cudaSuccess) return cuda_fail(e, "film copy");
    return check_last_frame(impl);
}

int pbrs_render_ids(const pbrs_scene *s, const pbrs_render_opts *o, uint32_t sample_index, uint32_t *out_inst, uint32_t *out_prim, float *out_t) {
    NEED(s && o, "render_ids: null argument");
    if (!s->impl.committed) return fail(PBRS_ERR_STATE, "render before pbrs_scene_commit");
    SceneImpl &impl = const_cast<pbrs_scene *>(s)->impl;
    cudaError_t e = cudaSetDevice(impl.device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
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
    rc = render_frame(impl, opts, tg, nullptr, nullptr);
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
    cudaError_t e = cudaSetDevice(impl.device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    NEED(o->world_size <= 1, "render_samples: single-rank only");
    uint32_t cw = o->crop_w ? o->crop_w : impl.cam.width, ch = o->crop_h ? o->crop_h : impl.cam.height;
    size_t n = (size_t)cw * ch * o->msaa * o->msaa * 3;
    DevBuf buf;
    int rc = buf.alloc(n * 4);
    if (rc < 0) return rc;
    RenderTargets tg;
    tg.samples = (float *)buf.p;
    rc = render_frame(impl, *o, tg, nullptr, st);
    if (rc < 0) return rc;
    if ((e = cudaMemcpy(out_rgb_samples, buf.p, n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "samples copy");
    return check_last_frame(impl);
}

}  // extern "C"
