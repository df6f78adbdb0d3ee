// Which feature of a kernel makes cudaOccupancyMaxActiveBlocksPerMultiprocessor answer 1 block per SM on B200?
// nvcc -gencode arch=compute_100a,code=sm_100a -o occ_probe occ_probe.cu && ./occ_probe   (the binary is not tracked)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(320, 2) k_plain(float* o) {
  extern __shared__ uint8_t sm[];
  sm[threadIdx.x] = threadIdx.x;
  __syncthreads();
  o[blockIdx.x * 320 + threadIdx.x] = sm[(threadIdx.x + 1) % 320];
}

__global__ void __launch_bounds__(320, 2) k_mbar(float* o) {
  extern __shared__ uint8_t sm[];
  uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  o[blockIdx.x * 320 + threadIdx.x] = sm[64 + threadIdx.x];
}

__global__ void __launch_bounds__(320, 2) k_pdl(float* o) {
  extern __shared__ uint8_t sm[];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  sm[threadIdx.x] = threadIdx.x;
  __syncthreads();
  o[blockIdx.x * 320 + threadIdx.x] = sm[(threadIdx.x + 1) % 320];
}

__global__ void __launch_bounds__(320, 2) k_tmem(float* o) {
  extern __shared__ uint8_t sm[];
  uint32_t slot = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  uint32_t t = *reinterpret_cast<volatile uint32_t*>(sm);
  o[blockIdx.x * 320 + threadIdx.x] = static_cast<float>(t);
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(64) : "memory");
}

// two CTAs per SM, each holding 64 TMEM columns while spinning until `expected` CTAs have arrived: proves (or
// disproves) co-residency on the hardware, bounded by a clock budget
__global__ void __launch_bounds__(320, 2) k_tmem_coresident(unsigned* counter, unsigned expected, int* result) {
  extern __shared__ uint8_t sm[];
  uint32_t slot = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  uint32_t t = *reinterpret_cast<volatile uint32_t*>(sm);
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1u);
    long long t0 = clock64();
    bool ok = false;
    while (clock64() - t0 < 200000000ll) {
      if (*reinterpret_cast<volatile unsigned*>(counter) >= expected) {
        ok = true;
        break;
      }
    }
    if (!ok) atomicExch(result, 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(64) : "memory");
}

template <typename K>
void probe(const char* name, K k) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  int a = 0, b = 0, c = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k, 320, 32768);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k, 320, 98304);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k, 128, 32768);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k);
  printf("%-10s regs %3d: blocks/SM at 320 thr 32K %d, 96K %d, 128 thr 32K %d\n", name, fa.numRegs, a, b, c);
}

int main() {
  probe("plain", k_plain);
  probe("mbarrier", k_mbar);
  probe("pdl", k_pdl);
  probe("tmem", k_tmem);
  probe("tmem-cores", k_tmem_coresident);
  unsigned* counter;
  int* result;
  cudaMalloc(&counter, 4);
  cudaMalloc(&result, 4);
  for (int smem : {32768, 98304}) {
    for (int grid : {148, 296}) {
      cudaMemset(counter, 0, 4);
      cudaMemset(result, 0, 4);
      k_tmem_coresident<<<grid, 320, smem>>>(counter, grid, result);
      cudaError_t e = cudaDeviceSynchronize();
      int r = -1;
      cudaMemcpy(&r, result, 4, cudaMemcpyDeviceToHost);
      printf("co-residency test: grid %d, %d B smem: %s (%s)\n", grid, smem, r == 0 ? "all CTAs resident together" : "TIMED OUT",
             cudaGetErrorString(e));
    }
  }
  return 0;
}
