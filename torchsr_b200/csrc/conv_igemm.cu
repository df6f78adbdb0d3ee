// Implicit-GEMM convolution / GEMM for sm_100a.
//
//   D[m, n] = sum_{tap, c} A[pixel(m) + tap, c] * B[tap][n][c]
//
// A (activations, NHWC bf16) is fetched by TMA in im2col mode straight from the unpadded tensor: one
// 128-pixel x block_k-channel box per (tap, channel chunk), halo pixels zero-filled by the TMA unit.
// B (packed weights, [tap][Cout][Cin] bf16, K-major) is fetched by tiled TMA. Both land in 128B/64B/32B
// swizzled shared memory and feed tcgen05.mma (M=128, N=block_n, K=16) with the fp32 accumulator in TMEM.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue (tcgen05.ld).
// The same kernel serves forward convs, stride-1 data gradients (flipped/transposed weight pack),
// stride-2 data gradients (one launch per output parity class) and the Linear layers (a_mode 1/2, split-K).
//
// Reference behaviour being replaced: every nn.Conv2d / nn.Linear call on the SRGAN/ESRGAN path
// (torchsr/srgan/generator.py:38-58, residual.py:27,64,67, discriminator.py:31-69 and the esrgan twins).
#include "conv_params.h"
#include "ptx.cuh"

namespace tsr {

namespace {

constexpr int kHeaderBytes = 10240;  // barriers + epilogue scratch, tiles start here (1024-aligned)
constexpr int kScratchOff = 1024;    // float scratch[4 warps][256 cols][2]

__device__ __forceinline__ void butterfly16(float (&v)[16], int lane, float& out) {
  // Sum each of the 16 per-lane values across the 32 lanes of the warp with 16 shuffles.
  // Result: lane l holds the total of column  ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1).
  float a[8], b[4], c[2], d;
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float send = u16 ? v[i] : v[i + 8];
    float keep = u16 ? v[i + 8] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float send = u8 ? a[i] : a[i + 4];
    float keep = u8 ? a[i + 4] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float send = u4 ? b[i] : b[i + 2];
    float keep = u4 ? b[i + 2] : b[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    float send = u2 ? c[0] : c[1];
    float keep = u2 ? c[1] : c[0];
    d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  out = d + __shfl_xor_sync(0xffffffffu, d, 1);
}

__device__ __forceinline__ int butterfly_col(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

__device__ __forceinline__ void load_bf16x16(const void* base, long long off, float (&z)[16]) {
  const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
  uint4 q0 = __ldg(p), q1 = __ldg(p + 1);
  unpack_bf16x2(q0.x, z[0], z[1]);
  unpack_bf16x2(q0.y, z[2], z[3]);
  unpack_bf16x2(q0.z, z[4], z[5]);
  unpack_bf16x2(q0.w, z[6], z[7]);
  unpack_bf16x2(q1.x, z[8], z[9]);
  unpack_bf16x2(q1.y, z[10], z[11]);
  unpack_bf16x2(q1.z, z[12], z[13]);
  unpack_bf16x2(q1.w, z[14], z[15]);
}

__device__ __forceinline__ void store_bf16x16(void* base, long long off, const float (&v)[16]) {
  uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off);
  uint4 q0, q1;
  q0.x = pack_bf16x2(v[0], v[1]);
  q0.y = pack_bf16x2(v[2], v[3]);
  q0.z = pack_bf16x2(v[4], v[5]);
  q0.w = pack_bf16x2(v[6], v[7]);
  q1.x = pack_bf16x2(v[8], v[9]);
  q1.y = pack_bf16x2(v[10], v[11]);
  q1.z = pack_bf16x2(v[12], v[13]);
  q1.w = pack_bf16x2(v[14], v[15]);
  p[0] = q0;
  p[1] = q1;
}

__device__ __forceinline__ void store_f32x16(void* base, long long off, const float (&v)[16]) {
  float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

}  // namespace

__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x;
  const int tile_n = blockIdx.y;
  const int split = blockIdx.z;

  const int stages = p.stages;
  const uint32_t a_bytes = kBlockM * p.block_k * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * p.block_k * 2;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
  const uint32_t bar_full = smem_base;                 // 8 x u64
  const uint32_t bar_empty = smem_base + 64;           // 8 x u64
  const uint32_t bar_tmem = smem_base + 128;           // u64
  const uint32_t tmem_slot = smem_base + 136;          // u32
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + 136);
  float* scratch = reinterpret_cast<float*>(smem_gen + kScratchOff);
  const uint32_t tiles = smem_base + kHeaderBytes;

  const int total_iters = p.num_taps * p.kc_per_tap;
  const int it_begin = split * p.iters_per_split;
  const int it_end = min(total_iters, it_begin + p.iters_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
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

  const int m0 = tile_m * kBlockM;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      int w0 = 0, h0 = 0, n0 = 0;
      if (p.a_mode == 0) {
        const int hw = p.Ho * p.Wo;
        n0 = m0 / hw;
        const int rem = m0 - n0 * hw;
        const int ho = rem / p.Wo;
        const int wo = rem - ho * p.Wo;
        h0 = ho * p.stride + p.lower_h;
        w0 = wo * p.stride + p.lower_w;
      }
      int idx = 0;
      for (int it = it_begin; it < it_end; ++it, ++idx) {
        const int s = idx % stages;
        const uint32_t ph = (idx / stages) & 1;
        if (!mbar_wait(bar_empty + 8 * s, ph ^ 1, p.epi.err, 1)) break;
        const int tap = it / p.kc_per_tap;
        const int kc = it - tap * p.kc_per_tap;
        const uint32_t a_dst = tiles + s * stage_bytes;
        const uint32_t b_dst = a_dst + a_bytes;
        const uint32_t full = bar_full + 8 * s;
        mbar_arrive_expect_tx(full, a_bytes + b_bytes);
        if (p.a_mode == 0) {
          const uint16_t off = p.tap_off[tap];
          tma_load_im2col_4d(a_dst, &p.tmA, full, p.a_c0 + kc * p.block_k, w0, h0, n0, off & 0xFF, off >> 8);
        } else if (p.a_mode == 1) {
          tma_load_2d(a_dst, &p.tmA, full, it * p.block_k, m0);
        } else {
          // MN-major A: two [block_k rows (K)] x [64 M-elements] boxes
          tma_load_2d(a_dst, &p.tmA, full, m0, it * p.block_k);
          tma_load_2d(a_dst + p.block_k * 128, &p.tmA, full, m0 + 64, it * p.block_k);
        }
        tma_load_2d(b_dst, &p.tmB, full, (p.a_mode == 0 ? kc : it) * p.block_k,
                    p.tap_wrow[tap] * p.b_rows_per_tap + tile_n * p.block_n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      const int row_bytes = p.block_k * 2;
      const uint32_t lt = layout_type_for_row_bytes(row_bytes);
      const uint32_t sbo = 8 * row_bytes;
      const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, p.a_mode == 2 ? 1 : 0, 0);
      int idx = 0;
      bool ok = true;
      for (int it = it_begin; it < it_end; ++it, ++idx) {
        const int s = idx % stages;
        const uint32_t ph = (idx / stages) & 1;
        if (!mbar_wait(bar_full + 8 * s, ph, p.epi.err, 2)) {
          ok = false;
          break;
        }
        tc_fence_after();
        const uint32_t a_src = tiles + s * stage_bytes;
        const uint32_t b_src = a_src + a_bytes;
        const int ksteps = p.block_k / 16;
        for (int k = 0; k < ksteps; ++k) {
          uint64_t adesc;
          if (p.a_mode == 2)
            adesc = make_smem_desc(a_src + k * 2048, /*LBO: next 64 M-elements*/ p.block_k * 128, 1024, 2);
          else
            adesc = make_smem_desc(a_src + k * 32, 16, sbo, lt);
          const uint64_t bdesc = make_smem_desc(b_src + k * 32, 16, sbo, lt);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (idx > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * s);
      }
      if (ok) umma_commit(bar_tmem);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const EpiParams& e = p.epi;
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool valid = m < p.M_total;
    int n = 0, ho = 0, wo = 0;
    if (p.a_mode == 0) {
      const int hw = p.Ho * p.Wo;
      n = m / hw;
      const int rem = m - n * hw;
      ho = rem / p.Wo;
      wo = rem - ho * p.Wo;
    }
    long long out_base = 0, aux_base = 0;
    if (e.out_mode == OUT_LINEAR) {
      out_base = n * e.os_n + ho * e.os_h + wo * e.os_w + e.out_ch_off;
    } else if (e.out_mode == OUT_UNSHUFFLE) {
      out_base = n * e.os_n + (ho >> 1) * e.os_h + (wo >> 1) * e.os_w + ((ho & 1) * 2 + (wo & 1)) * e.shuf_c +
                 e.out_ch_off;
    }
    aux_base = n * e.aux_n + ho * e.aux_h + wo * e.aux_w + e.aux_ch_off;
    if (p.a_mode != 0) {
      out_base = static_cast<long long>(m) * e.os_w + e.out_ch_off;
      aux_base = static_cast<long long>(m) * e.aux_w + e.aux_ch_off;
    }

    const float alpha = (e.prelu != nullptr) ? __ldg(e.prelu) : 0.f;
    float dalpha = 0.f;
    const bool ok = mbar_wait(bar_tmem, 0, e.err, 3);
    tc_fence_after();
    const int chunks = p.block_n / 16;
    if (ok) {
      for (int ch = 0; ch < chunks; ++ch) {
        uint32_t r[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ch * 16, r);
        tmem_ld_wait();
        const int col0 = tile_n * p.block_n + ch * 16;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * e.acc_scale;
        const bool st = valid && col0 < e.n_valid;
        if (e.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(e.bias + col0 + i);
        }
        if (e.bwd_z != nullptr && st) {
          float z[16];
          load_bf16x16(e.bwd_z, aux_base + col0, z);
          const float slope = (e.bwd_act == ACT_PRELU) ? alpha : (e.bwd_act == ACT_LEAKY ? e.leaky_slope : 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (z[i] <= 0.f) {
              dalpha += v[i] * z[i];
              v[i] *= slope;
            }
          }
        }
        if (e.res != nullptr && st) {
          float z[16];
          load_bf16x16(e.res, aux_base + col0, z);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += z[i];
        }
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        if (e.stats_partial != nullptr) {
          float sq[16], s1, s2;
#pragma unroll
          for (int i = 0; i < 16; ++i) sq[i] = v[i] * v[i];
          butterfly16(v, lane, s1);
          butterfly16(sq, lane, s2);
          if ((lane & 1) == 0) {
            const int c = ch * 16 + butterfly_col(lane);
            scratch[(q * 256 + c) * 2 + 0] = s1;
            scratch[(q * 256 + c) * 2 + 1] = s2;
          }
        }
        long long off;
        if (e.out_mode == OUT_SHUFFLE) {
          const int blk = col0 / e.shuf_c;
          off = n * e.os_n + (2 * ho + (blk >> 1)) * e.os_h + (2 * wo + (blk & 1)) * e.os_w + (col0 - blk * e.shuf_c) +
                e.out_ch_off;
        } else {
          off = out_base + col0;
        }
        if (st) {
          if (e.out_mode == OUT_GEMM_T_ATOMIC) {
            float* o = reinterpret_cast<float*>(e.out);
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(o + static_cast<long long>(col0 + i) * e.os_n + m, v[i]);
          } else {
            if (e.out_preact != nullptr) store_bf16x16(e.out_preact, off, v);
            if (e.act == ACT_PRELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * alpha;
            } else if (e.act == ACT_LEAKY) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * e.leaky_slope;
            } else if (e.act == ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (e.out_f32)
              store_f32x16(e.out, off, v);
            else
              store_bf16x16(e.out, off, v);
          }
        }
      }
    }
    // cross-warp reductions of the epilogue side products
    if (e.stats_partial != nullptr || e.dalpha_partial != nullptr) {
      if (e.dalpha_partial != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dalpha += __shfl_xor_sync(0xffffffffu, dalpha, o);
        if (lane == 0) scratch[4 * 256 * 2 + q] = dalpha;
      }
      named_bar_sync(1, 128);
      const int t = threadIdx.x - 64;  // 0..127
      if (e.stats_partial != nullptr && ok) {
        for (int c = t; c < p.block_n; c += 128) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            s1 += scratch[(w * 256 + c) * 2 + 0];
            s2 += scratch[(w * 256 + c) * 2 + 1];
          }
          float* dst = e.stats_partial + (static_cast<long long>(tile_m) * e.stats_ld + tile_n * p.block_n + c) * 2;
          dst[0] = s1;
          dst[1] = s2;
        }
      }
      if (e.dalpha_partial != nullptr && t == 0) {
        const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        e.dalpha_partial[cta] =
            scratch[4 * 256 * 2 + 0] + scratch[4 * 256 * 2 + 1] + scratch[4 * 256 * 2 + 2] + scratch[4 * 256 * 2 + 3];
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

size_t conv_igemm_smem_bytes(const ConvParams& p) {
  const uint32_t a_bytes = kBlockM * p.block_k * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * p.block_k * 2;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
  return 1024 + kHeaderBytes + static_cast<size_t>(p.stages) * stage_bytes;
}

cudaError_t launch_conv_igemm(const ConvParams& p, int tiles_n, int splits, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles_m = (p.M_total + kBlockM - 1) / kBlockM;
  dim3 grid(tiles_m, tiles_n, splits);
  conv_igemm_kernel<<<grid, kConvThreads, conv_igemm_smem_bytes(p), stream>>>(p);
  return cudaGetLastError();
}

}  // namespace tsr
