// Kernel launch helper: cudaLaunchKernelEx with the programmatic-dependent-launch attribute when `pdl` is set.
// `pdl` may only be set when the previous operation on `stream` is one of this library's kernels (all of which call
// pdl_wait(), see ptx.cuh): inside a captured CUDA graph the edge then becomes a programmatic dependency.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace tsr {

// cluster_y > 1 launches thread-block clusters of (1, cluster_y, 1) CTAs (gridDim.y must be a multiple of it).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                             int cluster_y, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_y > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = 1;
    at[n].val.clusterDim.y = static_cast<unsigned>(cluster_y);
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                            Args&&... args) {
  return launch_kc(kernel, grid, block, smem, stream, pdl, 1, std::forward<Args>(args)...);
}

}  // namespace tsr
