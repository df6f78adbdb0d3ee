// Weight-gradient of a convolution as a tcgen05 GEMM with K = pixels.
//
//   dW[co][tap][ci] += sum_{p in CTA's pixel range} X[pixel(p) + tap, ci] * dY[p, co]
//
// Both operands are NHWC (channel-contiguous), i.e. MN-major for this GEMM. X tiles arrive through TMA im2col
// (pix_per_stage pixels x chan_block channels per (tap, ci-block) "M block"), dY tiles through tiled TMA.
// blocks_per_m M blocks are stacked into one UMMA M=128 operand; groups_per_cta such groups share every dY
// tile and own block_n TMEM columns each. Partial sums over the pixel split are reduced with fp32
// red.global.add into a zero-initialised [Cout][tap][Cin] buffer (lanes = consecutive ci -> coalesced).
//
// Reference behaviour being replaced: autograd's convolution_backward (weight part) for every nn.Conv2d on
// the path, e.g. torchsr/srgan/residual.py:64,67.
#include "conv_params.h"
#include "ptx.cuh"

namespace tsr {

namespace {
constexpr int kWgHeader = 1024;
}

__global__ void __launch_bounds__(kConvThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int gset = blockIdx.x;   // which set of M groups
  const int tile_n = blockIdx.y;
  const int split = blockIdx.z;  // pixel range

  const int rowA = p.chan_block * 2;                 // bytes per pixel row in an A block
  const int rowB = p.dy_block * 2;
  const uint32_t blkA = p.pix_per_stage * rowA;      // bytes per A block per stage
  const uint32_t blkB = p.pix_per_stage * rowB;
  const int nb = p.block_n / p.dy_block;             // dY boxes per stage
  const int first_group = gset * p.groups_per_cta;
  const int first_block = first_group * p.blocks_per_m;
  const int cta_blocks = min(p.total_blocks - first_block, p.groups_per_cta * p.blocks_per_m);
  const int cta_groups = (cta_blocks + p.blocks_per_m - 1) / p.blocks_per_m;
  const uint32_t a_region = p.groups_per_cta * p.blocks_per_m * blkA;
  const uint32_t stage_bytes = (a_region + nb * blkB + 1023u) & ~1023u;

  const uint32_t bar_full = smem_base;
  const uint32_t bar_empty = smem_base + 64;
  const uint32_t bar_tmem = smem_base + 128;
  const uint32_t tmem_slot = smem_base + 136;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + 136);
  const uint32_t tiles = smem_base + kWgHeader;
  const int stages = p.stages;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmDy);
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tmem, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int pix_begin = split * p.stages_per_cta * p.pix_per_stage;
  int n_iters = p.stages_per_cta;
  {
    const int remaining = p.M_total - pix_begin;
    const int need = (remaining + p.pix_per_stage - 1) / p.pix_per_stage;
    n_iters = max(0, min(n_iters, need));
  }

  if (warp == 0) {
    if (lane == 0) {
      const int hw = p.Ho * p.Wo;
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        if (!mbar_wait(bar_empty + 8 * s, ph ^ 1, p.err, 11)) break;
        const int pix0 = pix_begin + it * p.pix_per_stage;
        const int n0 = pix0 / hw;
        const int rem = pix0 - n0 * hw;
        const int ho = rem / p.Wo;
        const int wo = rem - ho * p.Wo;
        const int h0 = ho * p.stride + p.lower_h;
        const int w0 = wo * p.stride + p.lower_w;
        const uint32_t a_dst = tiles + s * stage_bytes;
        const uint32_t b_dst = a_dst + a_region;
        const uint32_t full = bar_full + 8 * s;
        mbar_arrive_expect_tx(full, cta_blocks * blkA + nb * blkB);
        for (int j = 0; j < cta_blocks; ++j) {
          const int blk = first_block + j;
          const int tap = blk / p.cin_blocks;
          const int cib = blk - tap * p.cin_blocks;
          const uint16_t off = p.tap_off[tap];
          tma_load_im2col_4d(a_dst + j * blkA, &p.tmX, full, cib * p.chan_block, w0, h0, n0, off & 0xFF, off >> 8);
        }
        for (int j = 0; j < nb; ++j)
          tma_load_2d(b_dst + j * blkB, &p.tmDy, full, tile_n * p.block_n + j * p.dy_block, pix0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t ltA = layout_type_for_row_bytes(rowA);
      const uint32_t ltB = layout_type_for_row_bytes(rowB);
      const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, 1, 1);
      bool ok = true;
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        if (!mbar_wait(bar_full + 8 * s, ph, p.err, 12)) {
          ok = false;
          break;
        }
        tc_fence_after();
        const uint32_t a_src = tiles + s * stage_bytes;
        const uint32_t b_src = a_src + a_region;
        const int ksteps = p.pix_per_stage / 16;
        for (int g = 0; g < cta_groups; ++g) {
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t adesc =
                make_smem_desc(a_src + g * p.blocks_per_m * blkA + k * 16 * rowA, blkA, 8 * rowA, ltA);
            const uint64_t bdesc = make_smem_desc(b_src + k * 16 * rowB, blkB, 8 * rowB, ltB);
            umma_bf16(tmem_base + g * p.block_n, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_empty + 8 * s);
      }
      if (ok) umma_commit(bar_tmem);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int blk_in_group = row / p.chan_block;
    const int ci_in_blk = row - blk_in_group * p.chan_block;
    const bool ok = (n_iters > 0) && mbar_wait(bar_tmem, 0, p.err, 13);
    tc_fence_after();
    if (ok) {
      const int chunks = p.block_n / 16;
      for (int g = 0; g < cta_groups; ++g) {
        const int blk = first_block + g * p.blocks_per_m + blk_in_group;
        const bool row_ok = blk < p.total_blocks && blk < first_block + cta_blocks;
        const int tap = blk / p.cin_blocks;
        const int cib = blk - tap * p.cin_blocks;
        const int ci = cib * p.chan_block + ci_in_blk;
        for (int ch = 0; ch < chunks; ++ch) {
          uint32_t r[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * p.block_n + ch * 16, r);
          tmem_ld_wait();
          const int col0 = tile_n * p.block_n + ch * 16;
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int co = col0 + i;
              if (co < p.cout_valid)
                atomicAdd(p.out + (static_cast<long long>(co) * p.num_taps + tap) * p.cin_pad + ci,
                          __uint_as_float(r[i]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

size_t conv_wgrad_smem_bytes(const WgradParams& p) {
  const uint32_t blkA = p.pix_per_stage * p.chan_block * 2;
  const uint32_t blkB = p.pix_per_stage * p.dy_block * 2;
  const uint32_t a_region = p.groups_per_cta * p.blocks_per_m * blkA;
  const uint32_t stage_bytes = (a_region + (p.block_n / p.dy_block) * blkB + 1023u) & ~1023u;
  return 1024 + kWgHeader + static_cast<size_t>(p.stages) * stage_bytes;
}

cudaError_t launch_conv_wgrad(const WgradParams& p, int gsets, int tiles_n, int splits, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(gsets, tiles_n, splits);
  conv_wgrad_kernel<<<grid, kConvThreads, conv_wgrad_smem_bytes(p), stream>>>(p);
  return cudaGetLastError();
}

}  // namespace tsr
