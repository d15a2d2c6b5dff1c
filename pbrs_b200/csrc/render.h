// Host-side interface between the C ABI (api.cpp) and the wavefront driver (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "scene_host.h"

namespace pbrs {

// Device pointers the frame is written to; any may be null.
struct RenderTargets {
    float *film = nullptr;         // W*H*3, row-major, row 0 = top
    float *samples = nullptr;      // [crop_h][crop_w][spp][3]
    uint32_t *ids_inst = nullptr;  // with only_sample >= 0: primary-hit side channel, crop_w*crop_h each
    uint32_t *ids_prim = nullptr;
    float *ids_t = nullptr;
    int32_t only_sample = -1;
    // pbrs_render with a page-locked host film: the frame is cut into batches of whole tile rows and
    // every finished batch's rows are copied out on its own lane stream while the other lanes render
    // (render_frame sets host_copied when it did so; otherwise the caller copies the film at the end)
    float *host_film = nullptr;
    mutable bool host_copied = false;
};

struct Workspace;  // path buffers, queues and counters kept between calls (kernels.cu)
void workspace_free(Workspace *);

// Enqueues the whole frame on `stream`; synchronises only when `st` is given.
// One frame is in flight per replica at a time: the path workspace, counters and the cached graph
// belong to the replica.  The caller's current device is restored on return.
int render_frame(const SceneImpl &s, Replica &r, const pbrs_render_opts &o, const RenderTargets &tg, cudaStream_t stream, pbrs_stats *st);
// The 64x64 tiles of the render region that `rank` of `world` owns (tile split), in render order.
void owned_tiles(const SceneImpl &s, const pbrs_render_opts &o, std::vector<uint32_t> &tiles);

// After the frame's work has completed on the host side (the caller synchronised): fails loudly
// if a traversal stack overflowed during the last render_frame on this scene.
int check_last_frame(Replica &r);

}  // namespace pbrs
