// Kernel launch helper: cudaLaunchKernelEx with the programmatic-dependent-launch attribute when `pdl` is set.
// `pdl` may only be set when the previous operation on `stream` is one of this library's kernels (all of which call
// pdl_wait(), see ptx.cuh): inside a captured CUDA graph the edge then becomes a programmatic dependency.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace tsr {

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace tsr
