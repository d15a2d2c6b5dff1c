// Flattens the host scene into the HBM records of records.h and uploads them.
//
// Layout (DESIGN.md "Data layout"): every array is one contiguous allocation; records that a
// walk touches together are adjacent -- a mesh's inner nodes in preorder (the near child of a
// node is usually the next record), its triangles in leaf order, 48 bytes each with the
// primitive id and leaf-end flag in the pad lanes, so a leaf is a short sequential run.
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "render.h"

namespace pbrs {

struct DeviceArrays {
    std::vector<void *> allocs;
    size_t bytes = 0;
};

namespace {

template <class T>
int upload(DeviceArrays &d, const std::vector<T> &v, const T *&out) {
    out = nullptr;
    size_t n = v.size() * sizeof(T);
    void *p = nullptr;
    // keep every array non-null and 16-byte loadable even when empty
    cudaError_t e = cudaMalloc(&p, n + 64);
    if (e != cudaSuccess) { set_error(std::string("scene upload: ") + cudaGetErrorString(e)); return e == cudaErrorMemoryAllocation ? PBRS_ERR_OOM : PBRS_ERR_CUDA; }
    d.allocs.push_back(p);
    d.bytes += n;
    if (n) {
        e = cudaMemcpy(p, v.data(), n, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_error(std::string("scene upload: ") + cudaGetErrorString(e)); return PBRS_ERR_CUDA; }
    }
    out = reinterpret_cast<const T *>(p);
    return 0;
}

}  // namespace

static void replica_free(Replica &r) {
    int prev = -1;
    cudaGetDevice(&prev);
    if (r.device >= 0) cudaSetDevice(r.device);
    if (r.workspace) { workspace_free(r.workspace); r.workspace = nullptr; }
    if (r.film) { cudaFree(r.film); r.film = nullptr; r.film_bytes = 0; }
    if (r.dev) {
        for (void *p : r.dev->allocs) cudaFree(p);
        delete r.dev;
        r.dev = nullptr;
    }
    if (prev >= 0) cudaSetDevice(prev);
}

void device_free(SceneImpl &s) {
    replica_free(s);
    for (Replica *r : s.extra) { replica_free(*r); delete r; }
    s.extra.clear();
}

// Uploads the flattened records to the CURRENT device and points r.dscene at them.
static int upload_records(const SceneImpl &s, const FlatScene &f, Replica &r) {
    r.dev = new DeviceArrays();
    DeviceArrays &d = *r.dev;
    DeviceScene &ds = r.dscene;
    std::memset(&ds, 0, sizeof ds);
    int rc = 0;
    if ((rc = upload(d, s.tlas_nodes, ds.tlas_nodes)) < 0) return rc;
    if ((rc = upload(d, f.blas_nodes, ds.blas_nodes)) < 0) return rc;
    if ((rc = upload(d, f.tris, ds.tris)) < 0) return rc;
    if ((rc = upload(d, s.spheres, ds.spheres)) < 0) return rc;
    if ((rc = upload(d, s.simples, ds.simples)) < 0) return rc;
    if ((rc = upload(d, f.trav, ds.inst_trav)) < 0) return rc;
    if ((rc = upload(d, f.shade, ds.inst_shade)) < 0) return rc;
    if ((rc = upload(d, f.meshes, ds.meshes)) < 0) return rc;
    if ((rc = upload(d, f.tri_shade, ds.tri_shade)) < 0) return rc;
    if ((rc = upload(d, s.materials, ds.materials)) < 0) return rc;
    if ((rc = upload(d, f.textures, ds.textures)) < 0) return rc;
    if ((rc = upload(d, f.texels, ds.texels)) < 0) return rc;
    if ((rc = upload(d, f.perlin_vec, ds.perlin_vec)) < 0) return rc;
    if ((rc = upload(d, f.perlin_perm, ds.perlin_perm)) < 0) return rc;
    if ((rc = upload(d, s.delta_lights, ds.delta_lights)) < 0) return rc;
    if ((rc = upload(d, s.area_lights, ds.area_lights)) < 0) return rc;
    if ((rc = upload(d, f.blas_node_parent, ds.blas_node_parent)) < 0) return rc;
    if ((rc = upload(d, f.blas_leaf_parent, ds.blas_leaf_parent)) < 0) return rc;
    if ((rc = upload(d, f.tlas_node_parent, ds.tlas_node_parent)) < 0) return rc;
    if ((rc = upload(d, f.tlas_leaf_parent, ds.tlas_leaf_parent)) < 0) return rc;
    fill_scene_constants(s, f, ds);
    return 0;
}

int device_upload(SceneImpl &s) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (this back end has no CPU fallback)");
        return PBRS_ERR_NO_DEVICE;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("cudaGetDevice failed"); return PBRS_ERR_NO_DEVICE; }
    device_free(s);
    s.device = dev;
    FlatScene f;
    flatten_scene(s, f);
    int rc = upload_records(s, f, s);
    if (rc < 0) return rc;
    s.info.device_bytes = s.dev->bytes;
    return 0;
}

// A further copy of the committed scene on `device` (pbrs_render with num_gpus > 1).
int device_upload_replica(const SceneImpl &s, Replica &r, int device) {
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); set_error("replicate: cudaSetDevice failed"); return PBRS_ERR_NO_DEVICE; }
    r.device = device;
    FlatScene f;
    flatten_scene(s, f);
    int rc = upload_records(s, f, r);
    cudaSetDevice(prev);
    return rc;
}

}  // namespace pbrs
