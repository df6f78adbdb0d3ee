// Weight-gradient of a convolution as a tcgen05 GEMM with K = pixels.
//
//   dW[co][tap][ci] += sum_{p in CTA's pixel range} X[pixel(p) + tap, ci] * dY[p, co]
//
// Both operands are NHWC (channel-contiguous), i.e. MN-major for this GEMM. X tiles arrive through TMA im2col
// (pix_per_stage pixels x chan_block channels per (tap, ci-block) "M block"), dY tiles through tiled TMA.
// blocks_per_m M blocks are stacked into one UMMA M=128 operand; groups_per_cta such groups share every dY
// tile and own block_n TMEM columns each. Partial sums over the pixel split are reduced with 16-byte fp32 vector
// reductions (red.global.add.v4.f32) into a zero-initialised [tap][Cin][Cout_pad] buffer.
//
// Reference behaviour being replaced: autograd's convolution_backward (weight part) for every nn.Conv2d on
// the path, e.g. torchsr/srgan/residual.py:64,67.
#include "conv_params.h"
#include "launch.h"
#include "ptx.cuh"

namespace tsr {

namespace {
constexpr int kWgHeader = 1024;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgradThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
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

  pdl_sync();   // both operands (layer input, dY) come from earlier kernels of the stream
  const int pix_begin = split * p.stages_per_cta * p.pix_per_stage;
  int n_iters = p.stages_per_cta;
  {
    const int remaining = p.M_total - pix_begin;
    const int need = (remaining + p.pix_per_stage - 1) / p.pix_per_stage;
    n_iters = max(0, min(n_iters, need));
  }

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (incremental pixel coordinates)
      const int hw = p.Ho * p.Wo;
      int n0 = pix_begin / hw;
      int rem = pix_begin - n0 * hw;
      int ho = rem / p.Wo;
      int wo = rem - ho * p.Wo;
      // (tap, channel-block) of this CTA's first block; later blocks advance incrementally
      const int tap0 = first_block / p.cin_blocks;
      const int cib0 = first_block - tap0 * p.cin_blocks;
      const uint32_t tx = cta_blocks * blkA + nb * blkB;
      const int dy_col0 = tile_n * p.block_n;
      int s = 0;
      uint32_t ph = 1, dst = tiles;
      int pix0 = pix_begin;
      for (int it = 0; it < n_iters; ++it) {
        if (!mbar_wait(bar_empty + 8 * s, ph, p.err, 11)) break;
        const int h0 = ho * p.stride + p.lower_h;
        const int w0 = wo * p.stride + p.lower_w;
        const uint32_t full = bar_full + 8 * s;
        mbar_arrive_expect_tx(full, tx);
        int tap = tap0, cib = cib0;
        uint32_t a_dst = dst;
        for (int j = 0; j < cta_blocks; ++j) {
          const uint32_t off = p.tap_off[tap];
          tma_load_im2col_4d(a_dst, &p.tmX, full, cib * p.chan_block, w0, h0, n0, off & 0xFF, off >> 8);
          a_dst += blkA;
          if (++cib == p.cin_blocks) {
            cib = 0;
            ++tap;
          }
        }
        uint32_t b_dst = dst + a_region;
        for (int j = 0; j < nb; ++j) {
          tma_load_2d(b_dst, &p.tmDy, full, dy_col0 + j * p.dy_block, pix0);
          b_dst += blkB;
        }
        // advance the pixel cursor by pix_per_stage positions (W, then H, then N)
        pix0 += p.pix_per_stage;
        wo += p.pix_per_stage;
        while (wo >= p.Wo) {
          wo -= p.Wo;
          if (++ho == p.Ho) {
            ho = 0;
            ++n0;
          }
        }
        dst += stage_bytes;
        if (++s == stages) {
          s = 0;
          ph ^= 1;
          dst = tiles;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      const uint32_t ltA = layout_type_for_row_bytes(rowA);
      const uint32_t ltB = layout_type_for_row_bytes(rowB);
      const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, 1, 1);
      const uint64_t a0 = make_smem_desc(tiles, blkA, 8 * rowA, ltA);
      const uint64_t b0 = make_smem_desc(tiles + a_region, blkB, 8 * rowB, ltB);
      const uint32_t a_hi = static_cast<uint32_t>(a0 >> 32), b_hi = static_cast<uint32_t>(b0 >> 32);
      const uint32_t a_lo0 = static_cast<uint32_t>(a0), b_lo0 = static_cast<uint32_t>(b0);
      const uint32_t stage16 = stage_bytes >> 4, group16 = (p.blocks_per_m * blkA) >> 4;
      const uint32_t ka16 = (16 * rowA) >> 4, kb16 = (16 * rowB) >> 4;
      const int ksteps = p.pix_per_stage / 16;
      int s = 0;
      uint32_t ph = 0, soff = 0;
      bool ok = true;
      for (int it = 0; it < n_iters; ++it) {
        if (!mbar_wait(bar_full + 8 * s, ph, p.err, 12)) {
          ok = false;
          break;
        }
        tc_fence_after();
        const uint32_t acc = it > 0 ? 1u : 0u;
        uint32_t a_lo_g = a_lo0 + soff;
        uint32_t tm = tmem_base;
        for (int g = 0; g < cta_groups; ++g) {
          uint32_t a_lo = a_lo_g, b_lo = b_lo0 + soff;
          for (int k = 0; k < ksteps; ++k) {
            umma_bf16(tm, (static_cast<uint64_t>(a_hi) << 32) | a_lo, (static_cast<uint64_t>(b_hi) << 32) | b_lo, idesc,
                      (k > 0) ? 1u : acc);
            a_lo += ka16;
            b_lo += kb16;
          }
          a_lo_g += group16;
          tm += p.block_n;
        }
        umma_commit(bar_empty + 8 * s);
        soff += stage16;
        if (++s == stages) {
          s = 0;
          ph ^= 1;
          soff = 0;
        }
      }
      if (ok) umma_commit(bar_tmem);
    }
  } else {
    // -------------------------------------------------------------- epilogue: TMEM -> red.global.add.v4.f32
    // out[(tap * cin_pad + ci) * ld_out + co]: a thread owns one (tap, ci) row, its 16-column chunks are contiguous
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int blk_in_group = row / p.chan_block;
    const int ci_in_blk = row - blk_in_group * p.chan_block;
    const bool ok = (n_iters > 0) && mbar_wait(bar_tmem, 0, p.err, 13);
    tc_fence_after();
    if (ok) {
      const int chunks = p.block_n / 16;
      const int ld_out = p.cout_valid;   // padded output-channel count of the accumulator (multiple of 16)
      for (int g = 0; g < cta_groups; ++g) {
        const int blk = first_block + g * p.blocks_per_m + blk_in_group;
        const bool row_ok = blk < p.total_blocks && blk < first_block + cta_blocks;
        const int tap = blk / p.cin_blocks;
        const int cib = blk - tap * p.cin_blocks;
        const int ci = cib * p.chan_block + ci_in_blk;
        float* orow = p.out + (static_cast<long long>(tap) * p.cin_pad + ci) * ld_out + tile_n * p.block_n;
        for (int ch = 0; ch < chunks; ++ch) {
          uint32_t r[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * p.block_n + ch * 16, r);
          tmem_ld_wait();
          if (row_ok && tile_n * p.block_n + ch * 16 < ld_out) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              red_add_v4(orow + ch * 16 + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                         __uint_as_float(r[i + 3]));
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

cudaError_t launch_conv_wgrad(const WgradParams& p, int gsets, int tiles_n, int splits, cudaStream_t stream, bool pdl) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(gsets, tiles_n, splits);
  cudaError_t e = launch_k(conv_wgrad_kernel, grid, dim3(kWgradThreads), conv_wgrad_smem_bytes(p), stream, pdl, p);
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace tsr
