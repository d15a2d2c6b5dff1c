// Host-side interface between the C ABI (api.cpp) and the wavefront driver (kernels.cu).
#pragma once
#include <cuda_runtime.h>

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
};

struct Workspace;  // path buffers, queues and counters kept between calls (kernels.cu)
void workspace_free(Workspace *);

// Enqueues the whole frame on `stream`; synchronises only when `st` is given.
int render_frame(SceneImpl &s, const pbrs_render_opts &o, const RenderTargets &tg, cudaStream_t stream, pbrs_stats *st);

// After the frame's work has completed on the host side (the caller synchronised): fails loudly
// if a traversal stack overflowed during the last render_frame on this scene.
int check_last_frame(SceneImpl &s);

}  // namespace pbrs
