// Kernel launch helper: every kernel of the forward is launched with programmatic dependent launch
// (cudaLaunchAttributeProgrammaticStreamSerialization), optionally as a cluster.  A kernel launched this way
// may become resident while its predecessor in the stream is still draining; it does its set-up (barrier
// init, TMEM allocation, tensor-map prefetch) and then executes pdl_wait() (griddepcontrol.wait) before it
// touches any global memory, which returns once the predecessor grid has completed and its writes are
// visible.  VITDET_PDL=0 turns the attribute off (A/B measurements).
#pragma once

#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace vitdet {

inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VITDET_PDL"); v = (e && strcmp(e, "0") == 0) ? 0 : 1; }
    return v == 1;
}

// Opt-in to more than 48 KB of dynamic shared memory, per (kernel, device): the attribute is per device, and a process may
// drive more than one.  The configured size is remembered, and raised when a later launch of the same kernel needs more
// (a kernel whose shared-memory size depends on the model, e.g. the wide slot projection, may first run on a small one).
inline cudaError_t ensure_max_dynamic_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> done;      // (kernel, device) -> bytes already configured
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    int& have = done[std::make_pair(kernel, dev)];
    if (have >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace vitdet
