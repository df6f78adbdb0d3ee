// Implicit-GEMM convolution / GEMM for sm_100a.
//
//   D[m, n] = sum_{tap, c} A[pixel(m) + tap, c] * B[tap][n][c]
//
// A (activations, NHWC bf16) is fetched by TMA in im2col mode straight from the unpadded tensor: one
// 128-pixel x block_k-channel box per (tap, channel chunk), halo pixels zero-filled by the TMA unit.
// B (packed weights, [tap][Cout][Cin] bf16, K-major) is fetched by tiled TMA. Both land in 128B/64B/32B
// swizzled shared memory and feed tcgen05.mma (M=128, N=block_n, K=16) with the fp32 accumulator in TMEM.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue (tcgen05.ld; two warps
// per TMEM lane quadrant split the 16-column chunks).
// Instantiations: conv_body<A_MODE, PERS, FAST>. PERS = persistent weight-stationary CTAs (large-M im2col convs), FAST =
// lean producer / MMA loops for 64-channel K chunks without trace stamps or attribution hooks - the production path of
// every 3x3 / 9x9 conv whose input has a multiple of 64 channels (profiles/r01d_conv_attribution.md explains why the
// instruction count of those two single-warp loops sets the K-iteration rate).
// The same kernel serves forward convs, stride-1 data gradients (flipped/transposed weight pack),
// stride-2 data gradients (one launch per output parity class) and the Linear layers (a_mode 1/2, split-K).
// Epilogues (warps 2..9), selected per launch by the host (api.cu:build_conv, conv_params.h):
//   * generic chunk loop: bias / scale / residuals / activation forward or backward / PixelShuffle permutations /
//     BatchNorm column sums / nearest-x2 replication (out_rep2x) / OUT_GATHER_W (horizontal tap sums of the row-decomposed
//     9x9 output conv, one or two output rows per GEMM row);
//   * lean fused training BatchNorm forward and backward-apply paths (grid barrier, accumulator in registers);
//   * staged epilogue of the persistent kernels (one tcgen05.ld per warp, TMEM stage returned at once, arithmetic
//     instantiated per mode, 128B-swizzled staging buffer, cp.async.bulk.tensor store - or reducing store for in-place
//     residual blocks), on im2col tiles or on halo patches (3x3, or the 9x1 tap column of the output conv).
// Optional: activation multicast across clusters of two N tiles (TSR_CONV_CLUSTER=1, measured neutral).
//
// Reference behaviour being replaced: every nn.Conv2d / nn.Linear call on the SRGAN/ESRGAN path
// (torchsr/srgan/generator.py:38-58, residual.py:27,64,67, discriminator.py:31-69 and the esrgan twins).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "conv_params.h"
#include "launch.h"
#include "ptx.cuh"

namespace tsr {

namespace {

constexpr int kHeaderBytes = kConvHeaderBytes;  // barriers + epilogue scratch, tiles start here (1024-aligned)
constexpr int kScratchOff = 1024;    // float scratch[4 warps][256 cols][2]
constexpr int kQuadGroupOff = 9216 + 32;  // int qgrp[4]: BatchNorm statistics group of each TMEM lane quadrant's rows
constexpr int kColVecOff = 9216 + 64; // float colvec: bias[256], then BatchNorm scale[groups][256], shift[groups][256]
static_assert(kColVecOff + (1 + 2 * kMaxBnGroups) * 256 * 4 <= kHeaderBytes, "column vectors overflow the header");

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

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// OUT_GATHER_W for the generators' 9x9 Cout=3 output conv (EpiParams::gather_*), one 16-column chunk of a warp's 32
// accumulator rows (= 32 consecutive pixels, lane = pixel): column col = kw*3 + c is horizontal tap kw of channel c, and
//   out[pixel][c] = sum_kw acc[pixel + kw - 4][kw*3 + c].
// Lane L collects the taps that the warp's own rows hold for its pixel with one shuffle per column (a0); lanes 0..7 do
// the same for the 4 + 4 pixels just outside the warp's range (a1) - their remaining taps come from the neighbouring
// warp / tile, which adds them the same way. The caller sums the partials of the two column chunks and adds them to the
// zero-initialised fp32 NCHW output: at most two atomic adds ever meet on one element, so the result does not depend on
// their order. No shared-memory tile, no CTA barrier, and the 64-byte fp32 rows never travel to HBM.
template <int CH>
__device__ __forceinline__ void gather9x3(const float (&v)[16], int lane, int wo, int P, bool ext, int wcol, int Wo,
                                          float (&a0)[3], float (&a1)[3]) {
  constexpr int K = 9, GC = 3, PAD = 4;
#pragma unroll
  for (int c = 0; c < GC; ++c) a0[c] = a1[c] = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int col = CH * 16 + i;
    if (col < K * GC) {
      const int kw = col / GC, c = col % GC;
      const int src0 = lane + kw - PAD;
      const float t0 = __shfl_sync(0xffffffffu, v[i], src0 & 31);
      const int ws0 = wo + kw - PAD;
      if (src0 >= 0 && src0 < 32 && ws0 >= 0 && ws0 < Wo) a0[c] += t0;
      const int src1 = P + kw - PAD;
      const float t1 = __shfl_sync(0xffffffffu, v[i], src1 & 31);
      const int ws1 = wcol + kw - PAD;
      if (ext && src1 >= 0 && src1 < 32 && ws1 >= 0 && ws1 < Wo) a1[c] += t1;
    }
  }
}

// Grid-wide arrival barrier for the fused BatchNorm epilogues: `counter` is zero at launch, every CTA of the barrier's
// scope arrives once. Called by ONE thread after a CTA-level barrier; the fence makes the CTA's earlier global
// reductions (performed by other threads, ordered by that barrier) visible before the arrival, as cooperative-groups
// grid.sync() does. The host only selects this path for grids whose CTAs are all co-resident; the spin is bounded by
// the watchdog all the same (a timeout leaves a wrong tile and an error code, not a hung GPU).
__device__ __forceinline__ bool grid_arrive_and_wait(unsigned int* counter, unsigned int expected, int* err) {
  // release: orders this CTA's earlier reductions (made visible to this thread by the CTA barrier) before the arrival
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
  unsigned int seen;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
  if (seen < expected) {
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (clock64() - t0 > TSR_WATCHDOG_CYCLES) {
        if (err) atomicExch(err, 6);
        return false;
      }
    } while (seen < expected);
  }
  return true;
}

}  // namespace

// Shared-memory barrier block (offsets from the 1024-aligned base): full[8] @0, empty[8] @64, acc_full[2] @128,
// acc_empty[2] @144, b_full @160, TMEM slot @168, split-K flag @172.
//
// Persistent mode (p.persistent): gridDim.x CTAs walk the M tiles (tile_m = blockIdx.x, += gridDim.x). The whole weight
// operand of the CTA's N tile (num_taps * kc_per_tap K-chunks) is fetched ONCE into a resident region and only the
// activation tiles stream through the ring; two TMEM accumulator stages let the epilogue of tile j overlap the main
// loop of tile j+1. In the non-persistent mode gridDim.x == tiles_m, every CTA runs exactly one tile and both
// operands stream through the ring (weights of the first ring pass are fetched before the PDL wait when w_static).

// Warp 0 (all lanes run the loop so that every value stays warp-uniform; one elected lane issues): keeps the smem
// ring full. All per-iteration state is carried incrementally (no divisions inside the K loop).
template <int A_MODE, bool PERS, bool FAST>
__device__ __forceinline__ void producer_role(const ConvParams& p, uint32_t bar_full, uint32_t bar_empty, uint32_t bar_b,
                                              uint32_t b_res, uint32_t ring, int tile_n, int it_begin, int it_end,
                                              int tiles_m, long long* trace, int mc_rank) {
  constexpr bool pers = PERS;
  const int stages = p.stages;
  const uint32_t stage_bytes = p.stage_bytes, a_bytes = p.a_bytes, b_bytes = p.b_bytes;
  const uint32_t tx = pers ? a_bytes : a_bytes + b_bytes;
  const int kc_per_tap = p.kc_per_tap, block_k = p.block_k;
  const int b_row0 = tile_n * p.block_n;
  int pre = 0;
  auto load_resident_b = [&]() {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_b, static_cast<uint32_t>(it_end - it_begin) * b_bytes);
      uint32_t d = b_res;
      int ktap = 0, kkc = 0;
      for (int it = it_begin; it < it_end; ++it) {
        const int brow = (A_MODE == 0 ? p.tap_wrow[ktap] * p.b_rows_per_tap : 0) + b_row0;
        tma_load_2d(d, &p.tmB, bar_b, kkc * block_k, brow);
        d += b_bytes;
        if (++kkc == kc_per_tap && A_MODE == 0) {
          kkc = 0;
          ++ktap;
        }
      }
    }
    __syncwarp();
  };
  if (pers) {
    if (p.w_static) load_resident_b();
  } else {
    // Weight tiles of the first ring pass do not depend on the previous kernel of the stream (w_static): arm the
    // barriers and fetch them before the programmatic-launch wait, so only the activation loads follow it.
    pre = p.w_static ? min(stages, it_end - it_begin) : 0;
    if (elect_one()) {
      uint32_t d = ring + a_bytes;
      for (int j = 0; j < pre; ++j) {
        const int it = it_begin + j;
        int ktap = 0, kkc = it;
        if (A_MODE == 0) {
          ktap = it / kc_per_tap;
          kkc = it - ktap * kc_per_tap;
        }
        const int brow = (A_MODE == 0 ? p.tap_wrow[ktap] * p.b_rows_per_tap : kkc * p.b_chunk_rows) + b_row0;
        mbar_arrive_expect_tx(bar_full + 8 * j, tx);
        tma_load_2d(d, &p.tmB, bar_full + 8 * j, (A_MODE != 0 && p.b_chunk_rows) ? 0 : kkc * block_k, brow);
        d += stage_bytes;
      }
    }
    __syncwarp();
  }
  pdl_sync();
  if (pers && !p.w_static) load_resident_b();
  const bool leader = elect_one();   // the same lane every time; hoisted out of the FAST loop
  int s = 0;
  uint32_t ph = 1;  // parity to wait for on the empty barrier (first pass over the ring passes immediately)
  uint32_t dst = ring;
  bool first_tile = true;
  for (int tile_m = blockIdx.x; tile_m < tiles_m; tile_m += gridDim.x) {
    if (PERS && p.halo) {
      // one tiled load per 64-channel chunk: the (halo_th + 2) x halo_pw input patch of this tile, origin (-1, -1)
      const int n_img = tile_m / p.halo_tiles_per_img;
      const int t_img = tile_m - n_img * p.halo_tiles_per_img;
      const int t_y = t_img / p.halo_strips;
      const int hrow0 = t_y * p.halo_th + p.halo_lo_h;
      const int wcol0 = (t_img - t_y * p.halo_strips) * p.halo_sw + p.halo_lo_w;
      for (int kc = 0; kc < kc_per_tap; ++kc) {
        if (!mbar_wait(bar_empty + 8 * s, ph, p.epi.err, 1)) return;
        const uint32_t full = bar_full + 8 * s;
        if (FAST ? leader : elect_one()) {
          if (!FAST && (p.debug & 4)) {
            mbar_arrive(full);      // attribution run: no activation traffic, the MMAs read whatever the stage holds
          } else {
            mbar_arrive_expect_tx(full, p.a_bytes);
            tma_load_tile_4d(dst, &p.tmA, full, p.a_c0 + kc * block_k, wcol0, hrow0, n_img);
          }
        }
        if (!FAST) __syncwarp();
        dst += stage_bytes;
        if (++s == stages) {
          s = 0;
          ph ^= 1;
          dst = ring;
        }
      }
      continue;
    }
    // activation multicast (ConvParams::cluster_n): this CTA fetches the 64 pixels [m0 + 64 * rank, + 64) of the box
    const int m0 = tile_m * kBlockM + (mc_rank >= 0 ? 64 * mc_rank : 0);
    int w0 = 0, h0 = 0, n0 = 0;
    if (A_MODE == 0) {
      const int hw = p.Ho * p.Wo;
      n0 = m0 / hw;
      const int rem = m0 - n0 * hw;
      const int ho = rem / p.Wo;
      const int wo = rem - ho * p.Wo;
      h0 = ho * p.stride + p.lower_h;
      w0 = wo * p.stride_w + p.lower_w;
    }
    int tap = 0, kc = it_begin;
    if (A_MODE == 0) {
      tap = it_begin / kc_per_tap;
      kc = it_begin - tap * kc_per_tap;
    }
    if constexpr (FAST) {
      // Lean im2col loop (production path: 64-channel K chunks, no trace / attribution hooks). The K-iteration rate of
      // the small-N layers is set by the instruction latency of this single warp, not by TMA or the tensor core
      // (tools/trace_conv.py: with loads and UMMAs removed the loop still took ~500 clk per iteration). So: loop
      // invariants are pinned in registers, every operand of iteration i+1 (barrier addresses, destination, channel
      // coordinate, filter offsets) is prepared before the wait of iteration i+1, and the body is wait-arm-load.
      const uint32_t kcpt = in_reg(static_cast<uint32_t>(kc_per_tap)), nst = in_reg(static_cast<uint32_t>(stages));
      const uint32_t sbytes = in_reg(stage_bytes), txr = in_reg(tx), abytes = in_reg(a_bytes);
      const uint32_t b_tap_rows = in_reg(static_cast<uint32_t>(p.b_rows_per_tap));
      const int c_first = static_cast<int>(in_reg(static_cast<uint32_t>(p.a_c0)));
      const uint32_t mc_half = mc_rank > 0 ? (abytes >> 1) : 0u;
      int n_pre = first_tile ? pre : 0;
      uint32_t ukc = static_cast<uint32_t>(kc);
      // 64 channels per K chunk; the persistent instantiation also runs 32 (the one-tile kernel stays at 64: it is at its
      // 96-register budget)
      const int bk = PERS ? static_cast<int>(in_reg(static_cast<uint32_t>(block_k))) : 64;
      int cA = c_first + kc * bk, cB = kc * bk;
      uint32_t offs = p.tap_off[tap];
      int brow = static_cast<int>(p.tap_wrow[tap] * b_tap_rows) + b_row0;
      uint32_t full = bar_full + 8 * s, empty = bar_empty + 8 * s;
      for (int it = it_begin; it < it_end; ++it) {
        if (!mbar_wait(empty, ph, p.epi.err, 1)) return;
        if (leader) {
          if (PERS || n_pre <= 0) mbar_arrive_expect_tx(full, txr);
          if (!PERS && mc_rank >= 0)
            tma_load_im2col_4d_mc(dst + mc_half, &p.tmA2, full, cA, w0, h0, n0, offs & 0xFF, offs >> 8, 3);
          else
            tma_load_im2col_4d(dst, &p.tmA, full, cA, w0, h0, n0, offs & 0xFF, offs >> 8);
          if (!PERS && n_pre <= 0) tma_load_2d(dst + abytes, &p.tmB, full, cB, brow);
        }
        --n_pre;
        cA += bk;
        cB += bk;
        if (++ukc == kcpt) {
          ukc = 0;
          cA = c_first;
          cB = 0;
          ++tap;
          offs = p.tap_off[tap < kMaxTaps ? tap : kMaxTaps - 1];
          brow = static_cast<int>(p.tap_wrow[tap < kMaxTaps ? tap : kMaxTaps - 1] * b_tap_rows) + b_row0;
        }
        dst += sbytes;
        full += 8;
        empty += 8;
        if (static_cast<uint32_t>(++s) == nst) {
          s = 0;
          ph ^= 1;
          dst = ring;
          full = bar_full;
          empty = bar_empty;
        }
      }
      first_tile = false;
      if (!PERS) break;
      continue;
    }
    for (int it = it_begin; it < it_end; ++it) {
      if (!mbar_wait(bar_empty + 8 * s, ph, p.epi.err, 1)) return;
      const uint32_t full = bar_full + 8 * s;
      // arm the barrier unless the pre-issue above already did; weights go through the ring only when not resident
      const bool arm = !(first_tile && it - it_begin < pre);
      const bool load_b = arm && !pers;
      if (A_MODE == 0) {
        const uint32_t off = p.tap_off[tap];
        const int brow = p.tap_wrow[tap] * p.b_rows_per_tap + b_row0;
        if (elect_one()) {
          if (PERS && (p.debug & 4)) {
            mbar_arrive(full);      // attribution run (see the halo branch)
          } else {
            if (arm) mbar_arrive_expect_tx(full, tx);
            tma_load_im2col_4d(dst, &p.tmA, full, p.a_c0 + kc * block_k, w0, h0, n0, off & 0xFF, off >> 8);
            if (load_b) tma_load_2d(dst + a_bytes, &p.tmB, full, kc * block_k, brow);
          }
        }
        if (++kc == kc_per_tap) {
          kc = 0;
          ++tap;
        }
      } else if (A_MODE == 1) {
        if (elect_one()) {
          if (arm) mbar_arrive_expect_tx(full, tx);
          tma_load_2d(dst, &p.tmA, full, kc * block_k, m0);
          if (load_b) tma_load_2d(dst + a_bytes, &p.tmB, full, p.b_chunk_rows ? 0 : kc * block_k,
                                  kc * p.b_chunk_rows + b_row0);
        }
        ++kc;
      } else {
        // MN-major A: two [block_k rows (K)] x [64 M-elements] boxes
        if (elect_one()) {
          if (arm) mbar_arrive_expect_tx(full, tx);
          tma_load_2d(dst, &p.tmA, full, m0, kc * block_k);
          tma_load_2d(dst + block_k * 128, &p.tmA, full, m0 + 64, kc * block_k);
          if (load_b) tma_load_2d(dst + a_bytes, &p.tmB, full, p.b_chunk_rows ? 0 : kc * block_k,
                                  kc * p.b_chunk_rows + b_row0);
        }
        ++kc;
      }
      __syncwarp();
      if (trace && first_tile && it - it_begin < 16 && (threadIdx.x & 31) == 0) trace[8 + it - it_begin] = clock64();
      dst += stage_bytes;
      if (++s == stages) {
        s = 0;
        ph ^= 1;
        dst = ring;
      }
    }
    first_tile = false;
    if (!PERS) break;     // one tile per CTA: lets the compiler drop the loop-carried state
  }
}

// Warp 1 (all lanes loop, one elected lane issues): the UMMAs of every stage, slot recycling with tcgen05.commit.
template <int A_MODE, bool PERS, bool FAST>
__device__ __forceinline__ void mma_role(const ConvParams& p, uint32_t bar_full, uint32_t bar_empty, uint32_t bar_acc_full,
                                         uint32_t bar_acc_empty, uint32_t bar_b, uint32_t b_res, uint32_t ring,
                                         uint32_t tmem_base, int n_iters, int tiles_m, long long* trace, bool mc) {
  constexpr bool pers = PERS;
  const int stages = p.stages;
  const uint32_t ksteps = p.ksteps, idesc = p.idesc;
  const uint32_t stage16 = p.stage_bytes >> 4, a16 = p.a_bytes >> 4, b16 = p.b_bytes >> 4;
  // descriptor high words are loop invariant; the low word's start-address field advances in 16-byte units
  uint64_t a0, b0;
  uint32_t a_kadv;
  if (A_MODE == 2) {
    a0 = make_smem_desc(ring, p.block_k * 128, 1024, 2);
    a_kadv = 2048 >> 4;
  } else {
    a0 = make_smem_desc(ring, 16, p.sbo_bytes, p.layout_type);
    a_kadv = 32 >> 4;
  }
  b0 = make_smem_desc(pers ? b_res : ring, 16, p.sbo_bytes, p.layout_type);
  const uint32_t a_hi = static_cast<uint32_t>(a0 >> 32), b_hi = static_cast<uint32_t>(b0 >> 32);
  const uint32_t a_lo0 = static_cast<uint32_t>(a0), b_lo0 = static_cast<uint32_t>(b0) + (pers ? 0u : a16);
  if (pers) {
    if (!mbar_wait(bar_b, 0, p.epi.err, 4)) return;
    tc_fence_after();
  }
  const bool leader = elect_one();
  int s = 0, j = 0;
  uint32_t ph = 0, soff = 0;
  for (int tile_m = blockIdx.x; tile_m < tiles_m; tile_m += gridDim.x, ++j) {
    const int as = j & 1;
    // the epilogue must have drained this accumulator stage (first use of a stage passes immediately)
    if (!mbar_wait(bar_acc_empty + 8 * as, ((j >> 1) & 1) ^ 1, p.epi.err, 5)) return;
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * p.acc_cols;
    uint32_t acc = 0, boff = 0;
    if (PERS && FAST && p.halo) {
      // Lean halo loop: per 64-channel chunk nine 4-step UMMA groups (one per tap) read nine shifted windows of the
      // patch; window and weight-tile offsets are loop invariants.
      const uint32_t kcpt = static_cast<uint32_t>(p.kc_per_tap);
      const uint64_t a_desc0 = (static_cast<uint64_t>(a_hi) << 32) | a_lo0;
      const uint64_t b_desc0 = (static_cast<uint64_t>(b_hi) << 32) | b_lo0;
      for (uint32_t kc = 0; kc < kcpt; ++kc) {
        if (!mbar_wait(bar_full + 8 * s, ph, p.epi.err, 2)) return;
        tc_fence_after();
        if (leader) {
          const uint64_t ad = a_desc0 + static_cast<uint32_t>(s) * stage16;
          const uint64_t bd = b_desc0 + kc * b16;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t off = p.tap_off[t];
            const uint32_t win16 = ((off >> 8) * static_cast<uint32_t>(p.halo_pw) + (off & 0xFF)) * 8u;  // rows * 128 B / 16
            umma_bf16_x4(d_tmem, ad + win16, bd + static_cast<uint32_t>(t) * kcpt * b16, idesc, t == 0 ? acc : 1u);
          }
          umma_commit(bar_empty + 8 * s);
        }
        acc = 1;
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (leader) umma_commit(bar_acc_full + 8 * as);
      continue;
    }
    if (PERS && p.halo) {
      // per 64-channel chunk (one ring stage = one patch): nine taps = nine operand windows into the same patch
      const int kc_per_tap = p.kc_per_tap;
      for (int kc = 0; kc < kc_per_tap; ++kc) {
        if (!mbar_wait(bar_full + 8 * s, ph, p.epi.err, 2)) return;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t stage_addr = ring + static_cast<uint32_t>(s) * p.stage_bytes;
          for (int t = 0; t < p.num_taps; ++t) {
            const uint32_t off = p.tap_off[t];
            const uint32_t a_addr = stage_addr + ((off >> 8) * p.halo_pw + (off & 0xFF)) * 128u;
            // The window starts (a_addr / 128) % 8 rows into a 1024-byte swizzle atom. The tensor core derives the
            // 128B-swizzle phase from the absolute shared-memory address (as TMA did when it wrote the patch), so the
            // descriptor only needs the shifted start address; its base-offset field stays 0 (measured: setting it
            // to (addr >> 7) & 7 gives wrong results).
            uint32_t a_lo = (a_lo0 & ~0x3FFFu) | ((a_addr >> 4) & 0x3FFFu);
            const uint32_t a_hi_t = a_hi;
            uint32_t b_lo = b_lo0 + static_cast<uint32_t>(t * kc_per_tap + kc) * b16;
            for (uint32_t k = 0; k < ksteps; ++k) {
              if (!(p.debug & 8))
              umma_bf16(d_tmem, (static_cast<uint64_t>(a_hi_t) << 32) | a_lo, (static_cast<uint64_t>(b_hi) << 32) | b_lo,
                        idesc, acc);
              acc = 1;
              a_lo += a_kadv;
              b_lo += 2;
            }
          }
          umma_commit(bar_empty + 8 * s);
        }
        __syncwarp();
        acc = 1;
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (elect_one()) umma_commit(bar_acc_full + 8 * as);
      __syncwarp();
      continue;
    }
    if constexpr (FAST) {
      // Lean loop (64-element K chunks): one asm block issues the four UMMAs of the stage and the commit; descriptors
      // and barrier addresses of the next stage are advanced before its wait, loop invariants are pinned in registers.
      const uint32_t nst = in_reg(static_cast<uint32_t>(stages)), st16 = in_reg(stage16), b16r = in_reg(b16);
      const uint32_t idesc_r = in_reg(idesc);
      const bool k4 = PERS ? in_reg(ksteps) == 4u : true;
      const uint64_t a_desc0 = (static_cast<uint64_t>(a_hi) << 32) | a_lo0;
      const uint64_t b_desc0 = (static_cast<uint64_t>(b_hi) << 32) | b_lo0;
      uint64_t ad = a_desc0 + soff, bd = pers ? b_desc0 : b_desc0 + soff;
      uint32_t full = bar_full + 8 * s, empty = bar_empty + 8 * s;
      for (int it = 0; it < n_iters; ++it) {
        if (!mbar_wait(full, ph, p.epi.err, 2)) return;
        tc_fence_after();
        if (leader) {
          if (!PERS && mc)
            umma_bf16_x4_commit_mc(d_tmem, ad, bd, idesc_r, acc, empty, 3);   // frees the stage in both CTAs of the pair
          else if (k4)
            umma_bf16_x4_commit(d_tmem, ad, bd, idesc_r, acc, empty);
          else
            umma_bf16_x2_commit(d_tmem, ad, bd, idesc_r, acc, empty);         // 32-element K chunks
        }
        acc = 1;
        ad += st16;
        bd += pers ? b16r : st16;
        soff += st16;
        full += 8;
        empty += 8;
        if (static_cast<uint32_t>(++s) == nst) {
          s = 0;
          ph ^= 1;
          soff = 0;
          ad = a_desc0;
          if (!pers) bd = b_desc0;
          full = bar_full;
          empty = bar_empty;
        }
      }
      if (leader) umma_commit(bar_acc_full + 8 * as);
      if (!PERS) break;
      continue;
    }
    for (int it = 0; it < n_iters; ++it) {
      if (!mbar_wait(bar_full + 8 * s, ph, p.epi.err, 2)) return;
      tc_fence_after();
      if (trace && j == 0 && it < 16 && (threadIdx.x & 31) == 0) trace[24 + it] = clock64();
      if (elect_one()) {
        uint32_t a_lo = a_lo0 + soff, b_lo = b_lo0 + (pers ? boff : soff);
        uint32_t acc_k = acc;
        for (uint32_t k = 0; k < ksteps; ++k) {
          if (!(PERS && (p.debug & 8)))
          umma_bf16(d_tmem, (static_cast<uint64_t>(a_hi) << 32) | a_lo, (static_cast<uint64_t>(b_hi) << 32) | b_lo,
                    idesc, acc_k);
          acc_k = 1;
          a_lo += a_kadv;
          b_lo += 2;
        }
        umma_commit(bar_empty + 8 * s);
      }
      __syncwarp();
      acc = 1;
      boff += b16;
      soff += stage16;
      if (++s == stages) {
        s = 0;
        ph ^= 1;
        soff = 0;
      }
    }
    if (elect_one()) umma_commit(bar_acc_full + 8 * as);
    __syncwarp();
    if (trace && j == 0 && (threadIdx.x & 31) == 0) trace[4] = clock64();
    if (!PERS) break;
  }
}

// One CTA = one output tile (tile_m, tile_n) of conv `p`, K range = split `zsplit` of `nsplits`.
// FAST: production instantiation for im2col convs on 64-channel K chunks - no trace stamps, no attribution hooks,
// lean producer / MMA loops. The host selects it whenever the descriptor allows (launch_conv_igemm).
template <int A_MODE, bool PERS, bool FAST = false>
__device__ __forceinline__ void conv_body(const ConvParams& p, const int zsplit, const int nsplits) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_n = blockIdx.y;

  const uint32_t bar_full = smem_base;                 // 8 x u64
  const uint32_t bar_empty = smem_base + 64;           // 8 x u64
  const uint32_t bar_acc_full = smem_base + 128;       // 2 x u64
  const uint32_t bar_acc_empty = smem_base + 144;      // 2 x u64
  const uint32_t bar_b = smem_base + 160;              // u64 (resident weights)
  const uint32_t tmem_slot = smem_base + 168;          // u32
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + 168);
  float* scratch = reinterpret_cast<float*>(smem_gen + kScratchOff);
  const uint32_t b_res = smem_base + kHeaderBytes;     // resident weights (persistent mode) ...
  const uint32_t ring = b_res + p.b_res_bytes;         // ... then the operand ring
  const int tiles_m = (PERS && p.halo) ? p.halo_tiles_per_img * (p.M_total / (p.halo_H * p.halo_W))
                                       : (p.M_total + kBlockM - 1) / kBlockM;

  long long* trace = (!FAST && p.epi.trace)
                         ? p.epi.trace + 40ll * ((zsplit * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x)
                         : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = clock64();
  const int total_iters = p.num_taps * p.kc_per_tap;
  const int it_begin = zsplit * p.iters_per_split;
  const int it_end = min(total_iters, it_begin + p.iters_per_split);

  const bool mc = !PERS && FAST && A_MODE == 0 && p.cluster_n == 2;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(mc ? &p.tmA2 : &p.tmA);
    tma_prefetch_desc(&p.tmB);
    const int stages = p.stages;
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, mc ? 2 : 1);   // multicast stages are released by both CTAs' MMA warps
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, kConvThreads / 32 - 2);   // one arrival per epilogue warp
    }
    mbar_init(bar_b, 1);
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
  if (trace && threadIdx.x == 0) trace[1] = clock64();
  // the peer CTA multicasts into this CTA's ring and arrives on its barriers: both must have initialised them first
  if (mc) cluster_sync_all();

  if (warp == 0) {
    producer_role<A_MODE, PERS, FAST>(p, bar_full, bar_empty, bar_b, b_res, ring, tile_n, it_begin, it_end, tiles_m, trace,
                                      mc ? static_cast<int>(cluster_ctarank()) : -1);  // pdl_sync() inside
  } else if (warp == 1) {
    pdl_sync();
    mma_role<A_MODE, PERS, FAST>(p, bar_full, bar_empty, bar_acc_full, bar_acc_empty, bar_b, b_res, ring, tmem_base,
                     it_end - it_begin, tiles_m, trace, mc);
  } else {
    pdl_sync();   // the epilogue reads residuals / statistics buffers written by earlier kernels of the stream
    if (trace && threadIdx.x == 64) trace[2] = clock64();
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // TMEM lane quadrant q = warp % 4 (hardware rule); the two warps of a quadrant split the 16-column chunks.
    const EpiParams& e = p.epi;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int out_mode = e.out_mode, act = e.act, bwd_act = e.bwd_act, n_valid = e.n_valid, shuf_c = e.shuf_c;
    const float acc_scale = e.acc_scale, leaky = e.leaky_slope;
    const float* bias = e.bias;
    const void* bwd_z = e.bwd_z;
    const void* res = e.res;
    void* out = e.out;
    void* out_preact = e.out_preact;
    const bool want_stats = e.stats_partial != nullptr;
    const bool out_f32 = e.out_f32 != 0;
    const float alpha = (e.prelu != nullptr) ? __ldg(e.prelu) : 0.f;
    const float bslope = (bwd_act == ACT_PRELU) ? alpha : (bwd_act == ACT_LEAKY ? leaky : 0.f);
    const void* bnr_x = e.bnr_x;
    const int bnr_act = e.bnr_act;
    const float bnr_slope = (bnr_act == ACT_PRELU) ? __ldg(e.bnr_prelu) : (bnr_act == ACT_LEAKY ? leaky : 0.f);
    const int chunks = p.block_n >> 4;
    const int ch_begin = half ? (chunks + 1) >> 1 : 0;
    const int ch_end = half ? chunks : (chunks + 1) >> 1;
    const int colbase = tile_n * p.block_n;
    const int group_rows = e.group_rows;
    const int n_groups = group_rows > 0 ? 2 : 1;
    const int bnf = e.bnf_mode;
    const bool bnr_apply = e.bnr_apply != 0;
    // per-column vectors of this CTA's N tile -> shared memory, once (the chunk loop reads them as broadcasts)
    float* s_bias = reinterpret_cast<float*>(smem_gen + kColVecOff);
    float* s_sc = s_bias + 256;                   // [groups][256]
    float* s_sh = s_sc + kMaxBnGroups * 256;      // [groups][256]
    int* s_qgrp = reinterpret_cast<int*>(smem_gen + kQuadGroupOff);
    // Coefficients of the fused BatchNorm forward for this CTA's columns (every statistics group): training mode reads
    // the completed column sums (after the grid barrier), eval mode the running statistics. The CTA with blockIdx.x == 0
    // (first member of a grouped launch) publishes them for backward and updates the running statistics.
    auto bnf_coefficients = [&](bool publish) {
      const int et = threadIdx.x - 64;
      if (et < p.block_n) {
        const int c = colbase + et;
        const bool cv = c < e.bnf_c;
        const float gm = (cv && e.bnf_gamma) ? __ldg(e.bnf_gamma + c) : 1.f;
        const float bt = (cv && e.bnf_beta) ? __ldg(e.bnf_beta + c) : 0.f;
        float rm = 0.f, rv = 1.f;
        if (cv && e.bnf_rm != nullptr) {
          rm = e.bnf_rm[c];
          rv = e.bnf_rv[c];
        }
        for (int g = 0; g < n_groups; ++g) {
          float mean = rm, var = rv;
          if (bnf == 1) {
            const float inv_n = 1.f / static_cast<float>(e.bnf_count);
            const float* sp = e.stats_partial + (static_cast<long long>(g) * e.stats_ld + c) * 2;
            mean = cv ? __ldcg(sp) * inv_n : 0.f;
            var = cv ? fmaxf(__ldcg(sp + 1) * inv_n - mean * mean, 0.f) : 1.f;
          }
          const float invstd = rsqrtf(var + e.bnf_eps);
          const float sc = gm * invstd;
          const float sh = bt - mean * sc;
          s_sc[g * 256 + et] = sc;
          s_sh[g * 256 + et] = sh;
          if (publish && cv && bnf == 1) {
            if (e.bnf_coef != nullptr) {
              float* co = e.bnf_coef + static_cast<long long>(g) * 4 * e.bnf_c;
              co[0 * e.bnf_c + c] = sc;
              co[1 * e.bnf_c + c] = sh;
              co[2 * e.bnf_c + c] = mean;
              co[3 * e.bnf_c + c] = invstd;
            }
            const float n = static_cast<float>(e.bnf_count);
            const float unbiased = e.bnf_count > 1 ? var * n / (n - 1.f) : var;
            rm = (1.f - e.bnf_momentum) * rm + e.bnf_momentum * mean;
            rv = (1.f - e.bnf_momentum) * rv + e.bnf_momentum * unbiased;
          }
        }
        if (publish && cv && bnf == 1 && e.bnf_rm != nullptr) {
          e.bnf_rm[c] = rm;
          e.bnf_rv[c] = rv;
        }
      }
      if (publish && bnf == 1 && e.bnf_nbt != nullptr && blockIdx.y == 0 && et == 0) *e.bnf_nbt += n_groups;
      named_bar_sync(1, kConvThreads - 64);
    };
    {
      const int et = threadIdx.x - 64;
      if (et < p.block_n) {
        const int c = colbase + et;
        // PixelShuffle store: packed column c holds output channel 4*(c % shuf_c) + c / shuf_c; `bias` is the
        // parameter itself (OIHW channel order)
        const int bc = out_mode == OUT_SHUFFLE ? 4 * (c % shuf_c) + c / shuf_c : c;
        s_bias[et] = bias != nullptr ? __ldg(bias + bc) : 0.f;
        const bool hc = bnr_x != nullptr && e.bnr_coef != nullptr && c < e.bnr_c;
        if (bnf == 0) {
          for (int g = 0; g < n_groups; ++g) {
            const long long co = static_cast<long long>(g) * 4 * e.bnr_c + c;
            s_sc[g * 256 + et] = hc ? __ldg(e.bnr_coef + co) : 1.f;
            s_sh[g * 256 + et] = hc ? __ldg(e.bnr_coef + co + e.bnr_c) : 0.f;
          }
        }
      }
      named_bar_sync(1, kConvThreads - 64);
    }
    if (bnf == 2) bnf_coefficients(false);
    // shared memory behind the operand ring: staging buffers of the staged epilogue / the OUT_GATHER_W tile
    const uint32_t extra = ring + static_cast<uint32_t>(p.stages) * p.stage_bytes;
    bool staged_done = false;
    if constexpr (PERS) {   // (both instantiations: the 3-channel first layers run the K=32 non-FAST main loop)
      if (p.staged) {
        // ---- staged epilogue (conv_params.h): 64-column N tile, this warp owns 32 rows x 32 columns of it. One
        // tcgen05.ld + one wait per tile, the TMEM stage goes back to the MMA warp before any arithmetic, the bf16 tile
        // leaves through a 128B-swizzled staging buffer and ONE bulk tensor store per tile.
        staged_done = true;
        const int sw = p.halo ? p.halo_sw : 0;
        const bool has_bias = bias != nullptr;
        const float rs = e.res_scale;
        const float sl = act == ACT_PRELU ? alpha : (act == ACT_LEAKY ? leaky : 0.f);
        const int c_lo = half * 32;
        const long long aux_n = e.aux_n, aux_h = e.aux_h, aux_w = e.aux_w;
        const int aux_c = e.aux_ch_off + colbase + c_lo;
        const bool res_reduce = p.res_reduce != 0;     // the bulk store adds the tile to the residual already in `out`
        const bool res_cols_ok = res != nullptr && !res_reduce && colbase + c_lo < e.res_cols;
        const int row = q * 32 + lane;
        const bool shuf = out_mode == OUT_SHUFFLE;
        // tile -> (image, output pixel of this thread's accumulator row, its row in the staging buffer, box origin)
        struct TileGeom {
          int n, ho, wo, crow, w0, h0;
          bool valid;
        };
        auto tile_geom = [&](int tile_m) {
          TileGeom g;
          g.crow = row;
          g.w0 = g.h0 = 0;
          if (p.halo) {
            g.n = tile_m / p.halo_tiles_per_img;
            const int t_img = tile_m - g.n * p.halo_tiles_per_img;
            const int t_y = t_img / p.halo_strips;
            const int orow = row / p.halo_pw;
            const int pos = row - orow * p.halo_pw;
            g.w0 = (t_img - t_y * p.halo_strips) * sw;
            g.h0 = t_y * p.halo_th;
            g.wo = g.w0 + pos;
            g.ho = g.h0 + orow;
            g.valid = pos < sw && g.wo < p.halo_W && orow < p.halo_th && g.ho < p.halo_H;
            g.crow = orow * sw + pos;       // the two discarded positions of every patch row are squeezed out
          } else {
            const int m = tile_m * kBlockM + row;
            g.valid = m < p.M_total;
            const int hw = p.Ho * p.Wo;
            g.n = m / hw;
            const int rem = m - g.n * hw;
            g.ho = rem / p.Wo;
            g.wo = rem - g.ho * p.Wo;
          }
          return g;
        };
        // the residual row does not depend on the accumulator: it is fetched one whole tile ahead (registers), so its
        // L2 round trip never sits between two tiles of this warp
        auto load_res = [&](const TileGeom& g, uint4 (&rq)[4]) {
          const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(res) + g.n * aux_n +
                                                           g.ho * aux_h + g.wo * aux_w + aux_c);
#pragma unroll
          for (int k = 0; k < 4; ++k) rq[k] = __ldg(rp + k);
        };
        TileGeom tg = tile_geom(blockIdx.x);
        uint4 rq[4], rq_next[4];
        if (res_cols_ok && tg.valid && static_cast<int>(blockIdx.x) < tiles_m) load_res(tg, rq);
        int jj = 0;
        for (int tile_m = blockIdx.x; tile_m < tiles_m; tile_m += gridDim.x, ++jj) {
          const int as = jj & 1;
          const int n = tg.n, crow = tg.crow, w0 = tg.w0, h0 = tg.h0;
          const bool valid = tg.valid;
          const bool has_res = res_cols_ok && valid;
          const int tile_next = tile_m + static_cast<int>(gridDim.x);
          TileGeom tg_next = tg;
          if (tile_next < tiles_m) {
            tg_next = tile_geom(tile_next);
            if (res_cols_ok && tg_next.valid) load_res(tg_next, rq_next);
          }
          mbar_wait(bar_acc_full + 8 * as, (jj >> 1) & 1, e.err, 3);
          tc_fence_after();
          uint32_t r[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as) * p.acc_cols + c_lo, r);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);   // the MMA warp may start tile j+2 now
          uint32_t pk[16];
          // The arithmetic of the 32 columns, instantiated per (eval-BatchNorm | plain, activation kind, residual): with
          // the mode tests inside the element loop the section cost 1 500 - 2 900 clk per tile (tools/
          // microbench_epilogue.py: skipping it halved the launch) - uniform, but evaluated and branched on per element
          // by two warps per scheduler with nothing to hide the latency behind.
          auto math = [&](auto bn_c, auto act_c, auto res_c) {
            constexpr bool BN = decltype(bn_c)::value;
            constexpr int ACT = decltype(act_c)::value;      // 0 none, 1 slope (PReLU / LeakyReLU), 2 ReLU
            constexpr bool RES = decltype(res_c)::value;
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              float x[4], z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int k = 0; k < 4; ++k) x[k] = __uint_as_float(r[4 * i4 + k]);
              if (RES) {
                const uint4 qq = rq[i4 >> 1];
                unpack_bf16x2((i4 & 1) ? qq.z : qq.x, z[0], z[1]);
                unpack_bf16x2((i4 & 1) ? qq.w : qq.y, z[2], z[3]);
              }
              if (BN) {        // y = act(BN(acc)) + res
                const float4 sc4 = *reinterpret_cast<const float4*>(s_sc + c_lo + 4 * i4);
                const float4 sh4 = *reinterpret_cast<const float4*>(s_sh + c_lo + 4 * i4);
                x[0] = x[0] * sc4.x + sh4.x;
                x[1] = x[1] * sc4.y + sh4.y;
                x[2] = x[2] * sc4.z + sh4.z;
                x[3] = x[3] * sc4.w + sh4.w;
              } else {         // y = act((acc + bias) * acc_scale + res)
                const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c_lo + 4 * i4);   // zeros without a bias
                x[0] = (x[0] + b4.x) * acc_scale;
                x[1] = (x[1] + b4.y) * acc_scale;
                x[2] = (x[2] + b4.z) * acc_scale;
                x[3] = (x[3] + b4.w) * acc_scale;
                if (RES) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) x[k] += z[k] * rs;
                }
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (ACT == 2) x[k] = fmaxf(x[k], 0.f);
                if (ACT == 1) x[k] = x[k] > 0.f ? x[k] : x[k] * sl;
              }
              if (BN && RES) {
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] += z[k] * rs;
              }
              pk[2 * i4] = pack_bf16x2(x[0], x[1]);
              pk[2 * i4 + 1] = pack_bf16x2(x[2], x[3]);
            }
          };
          using T = std::true_type;
          using F = std::false_type;
          using A0 = std::integral_constant<int, 0>;
          using A1 = std::integral_constant<int, 1>;
          using A2 = std::integral_constant<int, 2>;
          const int act_kind = act == ACT_NONE ? 0 : (act == ACT_RELU ? 2 : 1);
          if (bnf == 2) {
            if (has_res) {
              if (act_kind == 0) math(T{}, A0{}, T{}); else if (act_kind == 1) math(T{}, A1{}, T{}); else math(T{}, A2{}, T{});
            } else {
              if (act_kind == 0) math(T{}, A0{}, F{}); else if (act_kind == 1) math(T{}, A1{}, F{}); else math(T{}, A2{}, F{});
            }
          } else {
            if (has_res) {
              if (act_kind == 0) math(F{}, A0{}, T{}); else if (act_kind == 1) math(F{}, A1{}, T{}); else math(F{}, A2{}, T{});
            } else {
              if (act_kind == 0) math(F{}, A0{}, F{}); else if (act_kind == 1) math(F{}, A1{}, F{}); else math(F{}, A2{}, F{});
            }
          }
          const uint32_t sbuf = extra + static_cast<uint32_t>(jj & 1) * 16384u;
          if (valid) {
            const uint32_t rowaddr = sbuf + static_cast<uint32_t>(crow) * 128u;
            const uint32_t sx = static_cast<uint32_t>(crow) & 7u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              sts128(rowaddr + (((static_cast<uint32_t>(half) * 4u + k) ^ sx) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2],
                     pk[4 * k + 3]);
          }
          fence_proxy_async_smem();
          // the store of tile j-1 (other buffer) must have left shared memory before any warp starts tile j+1
          if (threadIdx.x == 64) bulk_wait_group_read0();
          named_bar_sync(1, kConvThreads - 64);
          if (threadIdx.x == 64) {
            if (res_reduce) {
              if (p.halo)
                tma_reduce_add_4d(&p.tmO[0], sbuf, colbase, w0, h0, n);
              else
                tma_reduce_add_2d(&p.tmO[0], sbuf, colbase, tile_m * kBlockM);
            } else if (p.halo) {
              tma_store_4d(&p.tmO[shuf ? tile_n : 0], sbuf, shuf ? 0 : colbase, w0, h0, n);
            } else {
              tma_store_2d(&p.tmO[0], sbuf, colbase, tile_m * kBlockM);
            }
            bulk_commit_group();
          }
          tg = tg_next;
#pragma unroll
          for (int k = 0; k < 4; ++k) rq[k] = rq_next[k];
        }
        if (threadIdx.x == 64) bulk_wait_group0();
      }
    }
    int j = 0;
    for (int tile_m = blockIdx.x; !staged_done && tile_m < tiles_m; tile_m += gridDim.x, ++j) {
    const int m0 = tile_m * kBlockM;
    const int as = j & 1;          // accumulator stage of this tile
    const int row = q * 32 + lane;
    int m = m0 + row;
    bool valid = m < p.M_total;
    int n = 0, ho = 0, wo = 0;
    long long out_base, aux_base;
    if (A_MODE == 0) {
      if (PERS && p.halo) {
        // accumulator row -> (output row of the tile, position incl. the two discarded halo positions)
        n = tile_m / p.halo_tiles_per_img;
        const int t_img = tile_m - n * p.halo_tiles_per_img;
        const int t_y = t_img / p.halo_strips;
        const int orow = row / p.halo_pw;
        const int pos = row - orow * p.halo_pw;
        wo = (t_img - t_y * p.halo_strips) * p.halo_sw + pos;
        ho = t_y * p.halo_th + orow;
        valid = pos < p.halo_sw && wo < p.halo_W && orow < p.halo_th && ho < p.halo_H;
        m = (n * p.halo_H + ho) * p.halo_W + wo;
      } else {
        const int hw = p.Ho * p.Wo;
        n = m / hw;
        const int rem = m - n * hw;
        ho = rem / p.Wo;
        wo = rem - ho * p.Wo;
      }
      out_base = n * e.os_n + ho * e.os_h + wo * e.os_w + e.out_ch_off;
      if (e.out_mode == OUT_UNSHUFFLE)
        out_base = n * e.os_n + (ho >> 1) * e.os_h + (wo >> 1) * e.os_w + ((ho & 1) * 2 + (wo & 1)) * e.shuf_c +
                   e.out_ch_off;
      else if (e.out_mode == OUT_SHUFFLE)
        out_base = n * e.os_n + (2 * ho) * e.os_h + (2 * wo) * e.os_w + e.out_ch_off;
      aux_base = n * e.aux_n + ho * e.aux_h + wo * e.aux_w + e.aux_ch_off;
    } else {
      out_base = static_cast<long long>(m) * e.os_w + e.out_ch_off;
      aux_base = static_cast<long long>(m) * e.aux_w + e.aux_ch_off;
    }
    float dalpha = 0.f;
    // BatchNorm statistics group of this thread's row (all 32 rows of a warp share it: group_rows % 32 == 0)
    const int grp = (group_rows > 0 && m >= group_rows) ? 1 : 0;
    if (half == 0 && lane == 0) s_qgrp[q] = grp;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as) * p.acc_cols;
    // The auxiliary operands of the epilogue (residuals, activation-backward tensor, raw BatchNorm input) do not
    // depend on the accumulator: pull this thread's rows into L1 while the main loop runs, so that the dependent
    // global loads of the chunk loop below hit L1 instead of paying an L2 / HBM round trip per chunk.
    if (valid && out_mode != OUT_GEMM_T_ATOMIC) {
      const int c_lo = colbase + ch_begin * 16, c_hi = min(colbase + ch_end * 16, n_valid);
      const void* aux_ptrs[4] = {res, e.res2, bwd_z, bnr_x};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (aux_ptrs[a] == nullptr) continue;
        if (a < 2 && c_lo >= e.res_cols) continue;
        const char* rowp = reinterpret_cast<const char*>(aux_ptrs[a]) + (aux_base + c_lo) * 2;
        for (int b = 0; b < (c_hi - c_lo) * 2; b += 64) prefetch_l1(rowp + b);
      }
    }
    // y = act(v * scale + shift) + res * res_scale for one 16-column chunk of this thread's row (fused BatchNorm forward)
    auto bn_apply_store = [&](float (&v)[16], int ch, int col0) {
      const float* sc = s_sc + grp * 256 + ch * 16;
      const float* sh = s_sh + grp * 256 + ch * 16;
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = v[i] * sc[i] + sh[i];
      if (act == ACT_PRELU || act == ACT_LEAKY) {
        const float sl = act == ACT_PRELU ? alpha : leaky;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * sl;
      } else if (act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (res != nullptr && col0 < e.res_cols) {
        float z[16];
        load_bf16x16(res, aux_base + col0, z);
        const float rs = e.res_scale;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += z[i] * rs;
      }
      if (out_f32)
        store_f32x16(out, out_base + col0, v);
      else
        store_bf16x16(out, out_base + col0, v);
    };
    // ---- lean path of the fused training-mode BatchNorm (one tile per CTA, at most two 16-column chunks per warp: the
    // generator trunk and every narrow layer). The generic epilogue below re-reads its parameter block from the constant
    // bank inside every chunk and walks through every fusion option; tools/trace_fused.py measured ~1.0 us for each of
    // its two passes over ONE chunk, ~0.5 us to derive the coefficients and ~1.4 us around the grid barrier. Here every
    // invariant is in a register before the accumulator is ready, the BatchNorm parameters and the residual rows are
    // fetched while the main loop runs, and the accumulator stays in registers across the barrier.
    const int nch = ch_end - ch_begin;
    if (!PERS && bnf == 1 && nch <= 2 && nsplits == 1 && out_mode == OUT_LINEAR && !out_f32) {
      const int et = threadIdx.x - 64;
      const bool publish = blockIdx.x == 0 && blockIdx.z == 0;
      const bool col_thread = et < p.block_n;
      const int cc = colbase + et;
      const bool cv = col_thread && cc < e.bnf_c;
      float gm = 1.f, bt = 0.f, rm = 0.f, rv = 1.f;
      if (cv) {
        if (e.bnf_gamma) gm = __ldg(e.bnf_gamma + cc);
        if (e.bnf_beta) bt = __ldg(e.bnf_beta + cc);
        if (publish && e.bnf_rm != nullptr) {
          rm = e.bnf_rm[cc];
          rv = e.bnf_rv[cc];
        }
      }
      const float inv_n = 1.f / static_cast<float>(e.bnf_count), eps = e.bnf_eps, mom = e.bnf_momentum;
      const float cnt = static_cast<float>(e.bnf_count);
      float* const stats_g = e.stats_partial;
      const int stats_ld = e.stats_ld;
      unsigned int* const ctr = e.bnf_counter + blockIdx.y;
      const unsigned int expected = gridDim.x * gridDim.z;
      float* const coef_g = e.bnf_coef;
      const int bnf_c = e.bnf_c;
      const float rs = e.res_scale;
      const int res_cols = e.res_cols;
      const float sl = act == ACT_PRELU ? alpha : (act == ACT_LEAKY ? leaky : 0.f);
      // residual rows of this thread's chunks -> registers, now
      uint4 rq[2][2];
      bool has_r[2] = {false, false};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int col0 = colbase + (ch_begin + k) * 16;
        has_r[k] = k < nch && res != nullptr && valid && col0 < n_valid && col0 < res_cols;
        if (has_r[k]) {
          const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(res) + aux_base + col0);
          rq[k][0] = __ldg(rp);
          rq[k][1] = __ldg(rp + 1);
        }
      }
      const bool okl = mbar_wait(bar_acc_full + 8 * as, (j >> 1) & 1, e.err, 3);
      tc_fence_after();
      if (trace && threadIdx.x == 64) trace[5] = clock64();
      uint32_t r0[16], r1[16];
      tmem_ld16(taddr + ch_begin * 16, r0);
      if (nch == 2) tmem_ld16(taddr + (ch_begin + 1) * 16, r1);
      tmem_ld_wait();
      float v[2][16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[0][i] = (valid && okl) ? __uint_as_float(r0[i]) : 0.f;
        v[1][i] = (valid && okl && nch == 2) ? __uint_as_float(r1[i]) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k < nch) {
          float sq[16], s1, s2;
#pragma unroll
          for (int i = 0; i < 16; ++i) sq[i] = v[k][i] * v[k][i];
          butterfly16(v[k], lane, s1);
          butterfly16(sq, lane, s2);
          if ((lane & 1) == 0) {
            const int c = (ch_begin + k) * 16 + butterfly_col(lane);
            scratch[(q * 256 + c) * 2 + 0] = s1;
            scratch[(q * 256 + c) * 2 + 1] = s2;
          }
        }
      }
      if (trace && threadIdx.x == 64) trace[17] = clock64();
      named_bar_sync(1, kConvThreads - 64);
      for (int idx = et; idx < p.block_n * 2; idx += kConvThreads - 64) {
        const int c = idx >> 1, w = idx & 1;
        float sum0 = 0.f, sum1 = 0.f;
        bool any0 = false, any1 = false;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const float pv = scratch[(qq * 256 + c) * 2 + w];
          if (group_rows > 0 && s_qgrp[qq] != 0) {
            sum1 += pv;
            any1 = true;
          } else {
            sum0 += pv;
            any0 = true;
          }
        }
        if (any0) atomicAdd(stats_g + (colbase + c) * 2 + w, sum0);
        if (any1) atomicAdd(stats_g + (static_cast<long long>(stats_ld) + colbase + c) * 2 + w, sum1);
      }
      if (trace && threadIdx.x == 64) trace[18] = clock64();
      named_bar_sync(1, kConvThreads - 64);
      if (threadIdx.x == 64) grid_arrive_and_wait(ctr, expected, e.err);
      named_bar_sync(1, kConvThreads - 64);
      if (trace && threadIdx.x == 64) trace[19] = clock64();
      if (col_thread) {
        for (int g = 0; g < n_groups; ++g) {
          const float* sp = stats_g + (static_cast<long long>(g) * stats_ld + cc) * 2;
          const float mean = cv ? __ldcg(sp) * inv_n : 0.f;
          const float var = cv ? fmaxf(__ldcg(sp + 1) * inv_n - mean * mean, 0.f) : 1.f;
          const float invstd = rsqrtf(var + eps);
          const float sc = gm * invstd;
          const float sh = bt - mean * sc;
          s_sc[g * 256 + et] = sc;
          s_sh[g * 256 + et] = sh;
          if (publish && cv) {
            if (coef_g != nullptr) {
              float* co = coef_g + static_cast<long long>(g) * 4 * bnf_c;
              co[0 * bnf_c + cc] = sc;
              co[1 * bnf_c + cc] = sh;
              co[2 * bnf_c + cc] = mean;
              co[3 * bnf_c + cc] = invstd;
            }
            const float unbiased = cnt > 1.f ? var * cnt / (cnt - 1.f) : var;
            rm = (1.f - mom) * rm + mom * mean;
            rv = (1.f - mom) * rv + mom * unbiased;
          }
        }
        if (publish && cv && e.bnf_rm != nullptr) {
          e.bnf_rm[cc] = rm;
          e.bnf_rv[cc] = rv;
        }
      }
      if (publish && e.bnf_nbt != nullptr && blockIdx.y == 0 && et == 0) *e.bnf_nbt += n_groups;
      named_bar_sync(1, kConvThreads - 64);
      if (trace && threadIdx.x == 64) trace[20] = clock64();
      if (okl && valid) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int ch = ch_begin + k;
          const int col0 = colbase + ch * 16;
          if (k < nch && col0 < n_valid) {
            if (out_preact != nullptr) store_bf16x16(out_preact, out_base + col0, v[k]);   // raw conv output for backward
            const float* sc = s_sc + grp * 256 + ch * 16;
            const float* sh = s_sh + grp * 256 + ch * 16;
            float y[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float z = v[k][i] * sc[i] + sh[i];
              y[i] = (act == ACT_NONE || z > 0.f) ? z : z * sl;
            }
            if (has_r[k]) {
              float z[16];
              unpack_bf16x2(rq[k][0].x, z[0], z[1]);
              unpack_bf16x2(rq[k][0].y, z[2], z[3]);
              unpack_bf16x2(rq[k][0].z, z[4], z[5]);
              unpack_bf16x2(rq[k][0].w, z[6], z[7]);
              unpack_bf16x2(rq[k][1].x, z[8], z[9]);
              unpack_bf16x2(rq[k][1].y, z[10], z[11]);
              unpack_bf16x2(rq[k][1].z, z[12], z[13]);
              unpack_bf16x2(rq[k][1].w, z[14], z[15]);
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] += z[i] * rs;
            }
            store_bf16x16(out, out_base + col0, y);
          }
        }
      }
      if (trace && threadIdx.x == 64) trace[21] = clock64();
      if (trace && threadIdx.x == 64) trace[6] = clock64();
      break;
    }
    // ---- lean path of the fused BatchNorm-backward reduce + apply (same idea as the forward lean path above): dz and
    // the BatchNorm input x stay in registers across the grid barrier, every invariant is hoisted, x / residual rows and
    // the per-column parameters are fetched while the main loop runs. One chunk per warp (block_n <= 32: the generator
    // trunk) - two chunks of dz and x would not fit the 96-register budget of two CTAs per SM.
    if (!PERS && bnr_apply && nch == 1 && nsplits == 1 && out_mode == OUT_LINEAR && !out_f32 && bnr_x != nullptr &&
        e.bwd_z == nullptr && e.bias == nullptr && acc_scale == 1.f && e.res2 == nullptr) {
      const int et = threadIdx.x - 64;
      const bool publish = blockIdx.x == 0 && blockIdx.z == 0;
      const bool col_thread = et < p.block_n;
      const int cc = colbase + et;
      const bool cv = col_thread && cc < e.bnr_c;
      const int bc = e.bnr_c;
      float gm = 0.f, mu[kMaxBnGroups] = {0.f, 0.f}, is[kMaxBnGroups] = {1.f, 1.f};
      if (cv) {
        gm = __ldg(e.bnr_gamma + cc);
        for (int g = 0; g < n_groups; ++g) {
          mu[g] = __ldg(e.bnr_coef + (static_cast<long long>(g) * 4 + 2) * bc + cc);
          is[g] = __ldg(e.bnr_coef + (static_cast<long long>(g) * 4 + 3) * bc + cc);
        }
      }
      const float inv_m = 1.f / static_cast<float>(e.bnr_count);
      float* const stats_g = e.stats_partial;
      const int stats_ld = e.stats_ld;
      unsigned int* const ctr = e.bnf_counter + blockIdx.y;
      const unsigned int expected = gridDim.x * gridDim.z;
      const float rs = e.res_scale;
      const int res_cols = e.res_cols;
      float* const dalpha_g = e.dalpha_partial;
      void* const dx_out = e.bnr_dx;
      uint4 xq[1][2], rq[1][2];
      bool has_c[1] = {false}, has_r[1] = {false};
#pragma unroll
      for (int k = 0; k < 1; ++k) {
        const int col0 = colbase + (ch_begin + k) * 16;
        has_c[k] = k < nch && valid && col0 < n_valid;
        has_r[k] = has_c[k] && res != nullptr && col0 < res_cols;
        if (has_c[k]) {
          const uint4* xp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(bnr_x) + aux_base + col0);
          xq[k][0] = __ldg(xp);
          xq[k][1] = __ldg(xp + 1);
        }
        if (has_r[k]) {
          const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(res) + aux_base + col0);
          rq[k][0] = __ldg(rp);
          rq[k][1] = __ldg(rp + 1);
        }
      }
      const bool okl = mbar_wait(bar_acc_full + 8 * as, (j >> 1) & 1, e.err, 3);
      tc_fence_after();
      uint32_t r0[16];
      tmem_ld16(taddr + ch_begin * 16, r0);
      tmem_ld_wait();
      float dz[1][16], xr[1][16];
      auto unpack16 = [](const uint4 (&q)[2], float (&z)[16]) {
        unpack_bf16x2(q[0].x, z[0], z[1]);
        unpack_bf16x2(q[0].y, z[2], z[3]);
        unpack_bf16x2(q[0].z, z[4], z[5]);
        unpack_bf16x2(q[0].w, z[6], z[7]);
        unpack_bf16x2(q[1].x, z[8], z[9]);
        unpack_bf16x2(q[1].y, z[10], z[11]);
        unpack_bf16x2(q[1].z, z[12], z[13]);
        unpack_bf16x2(q[1].w, z[14], z[15]);
      };
#pragma unroll
      for (int k = 0; k < 1; ++k) {
        if (k < nch) {
          const int ch = ch_begin + k;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            dz[k][i] = (has_c[k] && okl) ? __uint_as_float(r0[i]) : 0.f;
            xr[k][i] = 0.f;
          }
          if (has_r[k]) {
            float z[16];
            unpack16(rq[k], z);
#pragma unroll
            for (int i = 0; i < 16; ++i) dz[k][i] += z[i] * rs;
          }
          if (has_c[k]) unpack16(xq[k], xr[k]);
          if (bnr_act != ACT_NONE) {
            const float* sc = s_sc + grp * 256 + ch * 16;
            const float* sh = s_sh + grp * 256 + ch * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float z = xr[k][i] * sc[i] + sh[i];
              if (z <= 0.f) {
                dalpha += dz[k][i] * z;
                dz[k][i] *= bnr_slope;
              }
            }
          }
          float sq[16], s1, s2;
#pragma unroll
          for (int i = 0; i < 16; ++i) sq[i] = dz[k][i] * xr[k][i];
          float t16[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) t16[i] = dz[k][i];
          butterfly16(t16, lane, s1);
          butterfly16(sq, lane, s2);
          if ((lane & 1) == 0) {
            const int c = ch * 16 + butterfly_col(lane);
            scratch[(q * 256 + c) * 2 + 0] = s1;
            scratch[(q * 256 + c) * 2 + 1] = s2;
          }
          // without an activation dz is the incoming gradient itself, which a residual block's skip path still reads
          if (bnr_act == ACT_NONE && has_c[k]) store_bf16x16(out, out_base + colbase + ch * 16, dz[k]);
        }
      }
      if (dalpha_g != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dalpha += __shfl_xor_sync(0xffffffffu, dalpha, o);
        if (lane == 0) scratch[4 * 256 * 2 + (warp - 2)] = dalpha;
      }
      named_bar_sync(1, kConvThreads - 64);
      for (int idx = et; idx < p.block_n * 2; idx += kConvThreads - 64) {
        const int c = idx >> 1, w = idx & 1;
        float sum0 = 0.f, sum1 = 0.f;
        bool any0 = false, any1 = false;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const float pv = scratch[(qq * 256 + c) * 2 + w];
          if (group_rows > 0 && s_qgrp[qq] != 0) {
            sum1 += pv;
            any1 = true;
          } else {
            sum0 += pv;
            any0 = true;
          }
        }
        if (any0) atomicAdd(stats_g + (colbase + c) * 2 + w, sum0);
        if (any1) atomicAdd(stats_g + (static_cast<long long>(stats_ld) + colbase + c) * 2 + w, sum1);
      }
      if (dalpha_g != nullptr && et == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += scratch[4 * 256 * 2 + w];
        atomicAdd(dalpha_g, tot);
      }
      named_bar_sync(1, kConvThreads - 64);
      if (threadIdx.x == 64) grid_arrive_and_wait(ctr, expected, e.err);
      named_bar_sync(1, kConvThreads - 64);
      float* cA = scratch;
      float* cB = scratch + kMaxBnGroups * 256;
      float* cC = scratch + 2 * kMaxBnGroups * 256;
      if (col_thread) {
        float dbeta = 0.f, dgamma = 0.f;
        for (int g = 0; g < n_groups; ++g) {
          float A = 0.f, Bx = 0.f, Cc = 0.f;
          if (cv) {
            const float* sp = stats_g + (static_cast<long long>(g) * stats_ld + cc) * 2;
            const float s1 = __ldcg(sp);
            const float s2 = is[g] * (__ldcg(sp + 1) - mu[g] * s1);
            const float c2 = s1 * inv_m, c3 = s2 * inv_m;
            A = gm * is[g];
            Bx = -A * c3 * is[g];
            Cc = -A * (c2 - mu[g] * is[g] * c3);
            dbeta += s1;
            dgamma += s2;
          }
          cA[g * 256 + et] = A;
          cB[g * 256 + et] = Bx;
          cC[g * 256 + et] = Cc;
        }
        if (publish && cv) {
          if (e.bnr_dbeta != nullptr) e.bnr_dbeta[cc] = dbeta;
          if (e.bnr_dgamma != nullptr) e.bnr_dgamma[cc] = dgamma;
        }
      }
      if (publish && blockIdx.y == 0 && et == 0 && e.bnr_dalpha != nullptr && dalpha_g != nullptr)
        *e.bnr_dalpha = __ldcg(dalpha_g);
      named_bar_sync(1, kConvThreads - 64);
#pragma unroll
      for (int k = 0; k < 1; ++k) {
        if (k < nch && has_c[k] && okl) {
          const int ch = ch_begin + k;
          const float* a = cA + grp * 256 + ch * 16;
          const float* b = cB + grp * 256 + ch * 16;
          const float* c = cC + grp * 256 + ch * 16;
          float dx[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) dx[i] = a[i] * dz[k][i] + b[i] * xr[k][i] + c[i];
          store_bf16x16(dx_out, out_base + colbase + ch * 16, dx);
        }
      }
      break;
    }
    const bool ok = mbar_wait(bar_acc_full + 8 * as, (j >> 1) & 1, e.err, 3);
    tc_fence_after();
    if (trace && j == 0 && threadIdx.x == 64) trace[5] = clock64();
    // ---- split-K: reduce the partial tiles through the fp32 workspace, the last CTA of the tile finalizes
    const bool split_ws = nsplits > 1 && out_mode != OUT_GEMM_T_ATOMIC;
    bool finalize = ok;
    if (split_ws) {
      float* wrow = e.ws + static_cast<long long>(m) * e.ws_ld + colbase;
      if (ok) {
        for (int ch = ch_begin; ch < ch_end; ++ch) {
          uint32_t r[16];
          tmem_ld16(taddr + ch * 16, r);     // .sync.aligned: every lane of the warp, valid row or not
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + ch * 16 + i), "r"(r[i]),
                           "r"(r[i + 1]), "r"(r[i + 2]), "r"(r[i + 3])
                           : "memory");
          }
        }
      }
      // release: the CTA's reductions are ordered before the counter bump by the barrier + one cumulative fence
      named_bar_sync(1, kConvThreads - 64);
      int* flag = reinterpret_cast<int*>(smem_gen + 172);
      if (threadIdx.x == 64) {
        __threadfence();
        int* cnt = e.tile_counters + blockIdx.y * tiles_m + tile_m;
        const int old = atomicAdd(cnt, 1);
        const int last = old == nsplits - 1;
        if (last) *cnt = 0;
        __threadfence();
        *flag = last;
      }
      named_bar_sync(1, kConvThreads - 64);
      finalize = ok && (*reinterpret_cast<volatile int*>(flag) != 0);
    }
    if (!FAST && (p.debug & 1)) finalize = false;
    if (finalize) {
      for (int ch = ch_begin; ch < ch_end; ++ch) {
        const int col0 = colbase + ch * 16;
        float v[16];
        if (split_ws) {
          // sums of all splits, read from L2 (every line is read once per launch and L1 starts each launch empty);
          // all loads of the chunk are issued before the zeroing stores - a store to an address waits for the
          // load of that address, so interleaving them would serialise on the L2 latency
          float4* wp = reinterpret_cast<float4*>(e.ws + static_cast<long long>(m) * e.ws_ld + col0);
          float4 q[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) q[i] = valid ? __ldcg(wp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[4 * i] = q[i].x; v[4 * i + 1] = q[i].y; v[4 * i + 2] = q[i].z; v[4 * i + 3] = q[i].w;
          }
          if (valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) __stcg(wp + i, make_float4(0.f, 0.f, 0.f, 0.f));
          }
        } else {
          uint32_t r[16];
          tmem_ld16(taddr + ch * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          if (PERS && p.halo && !valid) {
            // halo / padding rows of the tile were computed from whatever the patch buffer held: keep their
            // (possibly non-finite) values out of the side reductions below
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
        }
        if (trace && threadIdx.x == 64 && ch < 4) trace[32 + 2 * ch] = clock64();
        if (out_mode == OUT_GATHER_W) {
          if (!valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
          // pixels just outside the warp's rows: lanes 0..3 -> first - 4 .. - 1, lanes 4..7 -> first + 32 .. + 35
          const int P = lane < 4 ? lane - 4 : 28 + lane;
          // (m - lane = linear index of the warp's first pixel: the warp's 32 rows are 32 consecutive pixels of one
          // image row, in im2col tiles and - patch width a multiple of 32 - in halo tiles alike)
          const long long mo = static_cast<long long>(m) - lane + P;
          const bool ext = lane < 8 && mo >= 0 && mo < p.M_total;
          int rowid = 0, wcol = 0;
          if (ext) {
            rowid = static_cast<int>(mo / p.Wo);
            wcol = static_cast<int>(mo - static_cast<long long>(rowid) * p.Wo);
          }
          float a0[3], a1[3];
          const int cl = e.gather_rows == 2 ? ((ch - ch_begin) & 1) : ch;     // chunk inside the 32 columns of an output row
          if (cl == 0)
            gather9x3<0>(v, lane, wo, P, ext, wcol, p.Wo, a0, a1);
          else
            gather9x3<1>(v, lane, wo, P, ext, wcol, p.Wo, a0, a1);
          float* const o = reinterpret_cast<float*>(out);
          if (e.gather_rows == 2) {
            // two output rows per GEMM row: this warp holds BOTH column chunks of output row 2*ho + half; the first
            // chunk's partial sums wait in registers for the second
            // (parked in this thread's own slot of the reduction scratch rather than in registers: six registers live
            // across the chunk loop made the one-tile instantiation spill)
            float* const park = scratch + (half * kBlockM + row) * 6;
            if (cl == 0) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                park[c] = a0[c];
                park[3 + c] = a1[c];
              }
            } else {
              float g2a0[3], g2a1[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                g2a0[c] = park[c];
                g2a1[c] = park[3 + c];
              }
              const int Hout = 2 * p.Ho;
              const long long cs = static_cast<long long>(Hout) * p.Wo;
              if (valid) {
                const long long own = ((static_cast<long long>(n) * 3) * Hout + 2 * ho + half) * p.Wo + wo;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                  atomicAdd(o + own + c * cs, a0[c] + g2a0[c] + (e.gather_bias != nullptr ? __ldg(e.gather_bias + c) : 0.f));
              }
              if (ext) {
                const int nn = rowid / p.Ho, hh = rowid - nn * p.Ho;
                const long long oe = ((static_cast<long long>(nn) * 3) * Hout + 2 * hh + half) * p.Wo + wcol;
#pragma unroll
                for (int c = 0; c < 3; ++c) atomicAdd(o + oe + c * cs, a1[c] + g2a1[c]);
              }
            }
            continue;
          }
          // the two warps of a lane quadrant hold the two column chunks of the same rows: the upper half hands its
          // partial sums over through shared memory (double-buffered across the tiles of a persistent CTA)
          float* pbuf = scratch + ((j & 1) * kBlockM + row) * 6;
          // (barrier ids spelled out: a run-time id makes ptxas reserve all 16 hardware barriers for the CTA)
          auto pair_sync = [&]() {
            if (q == 0) named_bar_sync(2, 64);
            else if (q == 1) named_bar_sync(3, 64);
            else if (q == 2) named_bar_sync(4, 64);
            else named_bar_sync(5, 64);
          };
          if (half == 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              pbuf[c] = a0[c];
              pbuf[3 + c] = a1[c];
            }
            pair_sync();
          } else {
            pair_sync();
            const long long cs = static_cast<long long>(p.Ho) * p.Wo;
            if (valid) {
              const long long own = ((static_cast<long long>(n) * 3) * p.Ho + ho) * p.Wo + wo;
#pragma unroll
              for (int c = 0; c < 3; ++c)
                atomicAdd(o + own + c * cs, a0[c] + pbuf[c] + (e.gather_bias != nullptr ? __ldg(e.gather_bias + c) : 0.f));
            }
            if (ext) {
              const int nn = rowid / p.Ho, hh = rowid - nn * p.Ho;
              const long long oe = ((static_cast<long long>(nn) * 3) * p.Ho + hh) * p.Wo + wcol;
#pragma unroll
              for (int c = 0; c < 3; ++c) atomicAdd(o + oe + c * cs, a1[c] + pbuf[3 + c]);
            }
          }
          continue;
        }
        const bool st = valid && col0 < n_valid && (FAST || !(p.debug & 2));
        if (bnf == 2) {             // eval-mode BatchNorm folded into this pass: coefficients are known up front
          if (st) bn_apply_store(v, ch, col0);
          continue;
        }
        if (bias != nullptr) {
          const float4* bp = reinterpret_cast<const float4*>(s_bias + ch * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = bp[i];
            v[4 * i] += b4.x;
            v[4 * i + 1] += b4.y;
            v[4 * i + 2] += b4.z;
            v[4 * i + 3] += b4.w;
          }
        }
        if (acc_scale != 1.f) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= acc_scale;
        }
        if (bwd_z != nullptr && st) {
          float z[16];
          load_bf16x16(bwd_z, aux_base + col0, z);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (z[i] <= 0.f) {
              dalpha += v[i] * z[i];
              v[i] *= bslope;
            }
          }
        }
        if (res != nullptr && st && col0 < e.res_cols && bnf == 0) {
          float z[16];
          load_bf16x16(res, aux_base + col0, z);
          const float rs = e.res_scale;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += z[i] * rs;
          if (e.res2 != nullptr) {
            load_bf16x16(e.res2, aux_base + col0, z);
            const float rs2 = e.res2_scale;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += z[i] * rs2;
          }
        }
        float xr[16];
        if (bnr_x != nullptr) {
          if (st) {
            load_bf16x16(bnr_x, aux_base + col0, xr);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) xr[i] = 0.f;
          }
          if (bnr_act != ACT_NONE) {
            const float* sc = s_sc + grp * 256 + ch * 16;
            const float* sh = s_sh + grp * 256 + ch * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float z = xr[i] * sc[i] + sh[i];
              if (z <= 0.f) {
                dalpha += v[i] * z;
                v[i] *= bnr_slope;
              }
            }
          }
        }
        if (want_stats) {
          if (!valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
          float sq[16], s1, s2;
#pragma unroll
          for (int i = 0; i < 16; ++i) sq[i] = v[i] * (bnr_x != nullptr ? xr[i] : v[i]);
          butterfly16(v, lane, s1);
          butterfly16(sq, lane, s2);
          if ((lane & 1) == 0) {
            const int c = ch * 16 + butterfly_col(lane);
            scratch[(q * 256 + c) * 2 + 0] = s1;
            scratch[(q * 256 + c) * 2 + 1] = s2;
          }
        }
        if (bnr_apply) {
          // fused BatchNorm-backward apply: park dz in the accumulator (warp-collective, every lane) - dx is formed from
          // it after the grid barrier below
          uint32_t pz[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pz[i] = __float_as_uint(st ? v[i] : 0.f);
          tmem_st16(taddr + ch * 16, pz);
          // with an activation nothing else reads dz; without one dz == the incoming gradient, which a residual
          // block's skip path still needs: it is stored to `out` as usual, dx goes to bnr_dx
          if (bnr_act != ACT_NONE) continue;
        }
        if (st) {
          long long off = out_base + col0;
          if (out_mode == OUT_SHUFFLE) {
            const int blk = col0 / shuf_c;
            off = out_base + (blk >> 1) * e.os_h + (blk & 1) * e.os_w + (col0 - blk * shuf_c);
          }
          if (out_mode == OUT_GEMM_T_ATOMIC) {
            float* o = reinterpret_cast<float*>(out);
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(o + static_cast<long long>(col0 + i) * e.os_n + m, v[i]);
          } else {
            if (bnf != 0) continue;     // raw and normalised values are stored after the grid barrier (below)
            if (out_preact != nullptr) store_bf16x16(out_preact, off, v);
            if (act == ACT_PRELU || act == ACT_LEAKY) {
              const float sl = act == ACT_PRELU ? alpha : leaky;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * sl;
            } else if (act == ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (out_f32)
              store_f32x16(out, off, v);
            else
              store_bf16x16(out, off, v);
            if (A_MODE == 0 && e.out_rep2x != nullptr) {
              // nearest x2 of the result (EpiParams::out_rep2x): the same 16 values to the 2 x 2 fine-grid positions
              const long long rb = n * e.rep_n + (2 * ho) * e.rep_h + (2 * wo) * e.rep_w + e.rep_ch_off + col0;
              store_bf16x16(e.out_rep2x, rb, v);
              store_bf16x16(e.out_rep2x, rb + e.rep_w, v);
              store_bf16x16(e.out_rep2x, rb + e.rep_h, v);
              store_bf16x16(e.out_rep2x, rb + e.rep_h + e.rep_w, v);
            }
          }
        }
        if (trace && threadIdx.x == 64 && ch < 4) trace[33 + 2 * ch] = clock64();
      }
    }
    if (trace && j == 0 && threadIdx.x == 64) trace[17] = clock64();
    // every tcgen05.ld of this tile has completed (tmem_ld_wait above): hand the accumulator stage back to the MMA warp
    // (the fused training BatchNorm / BatchNorm-backward apply read the accumulator once more after the grid barrier and
    // arrive there)
    if (bnr_apply) tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (bnf != 1 && !bnr_apply && lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
    // cross-warp reductions of the epilogue side products -> one red.global.add per column / per tile
    if (want_stats || e.dalpha_partial != nullptr) {
      if (e.dalpha_partial != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dalpha += __shfl_xor_sync(0xffffffffu, dalpha, o);
        if (lane == 0) scratch[4 * 256 * 2 + (warp - 2)] = dalpha;
      }
      named_bar_sync(1, kConvThreads - 64);
      const int t = threadIdx.x - 64;  // 0..255
      if (want_stats && finalize) {
        // thread t -> (column t>>1, statistic t&1) for block_n <= 128; two passes for wider tiles. The four lane
        // quadrants' partial sums go to the statistics group their rows belong to (one group unless group_rows > 0).
        for (int idx = t; idx < p.block_n * 2; idx += kConvThreads - 64) {
          const int c = idx >> 1, w = idx & 1;
          float sum0 = 0.f, sum1 = 0.f;
          bool any0 = false, any1 = false;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const float pv = scratch[(qq * 256 + c) * 2 + w];
            if (group_rows > 0 && s_qgrp[qq] != 0) {
              sum1 += pv;
              any1 = true;
            } else {
              sum0 += pv;
              any0 = true;
            }
          }
          if (any0) atomicAdd(e.stats_partial + (colbase + c) * 2 + w, sum0);
          if (any1) atomicAdd(e.stats_partial + (static_cast<long long>(e.stats_ld) + colbase + c) * 2 + w, sum1);
        }
      }
      if (e.dalpha_partial != nullptr && t == 0 && finalize) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += scratch[4 * 256 * 2 + w];
        atomicAdd(e.dalpha_partial, tot);
      }
      // the scratch slots are reused by the next tile of a persistent CTA
      if (tile_m + static_cast<int>(gridDim.x) < tiles_m) named_bar_sync(1, kConvThreads - 64);
    }
    if (trace && j == 0 && threadIdx.x == 64) trace[18] = clock64();
    if (bnf == 1) {
      // ---- fused training-mode BatchNorm: wait until every CTA of this N tile has added its column sums, derive the
      // coefficients, then normalise + activate (+ residual) straight from the accumulator
      named_bar_sync(1, kConvThreads - 64);
      if (threadIdx.x == 64) grid_arrive_and_wait(e.bnf_counter + blockIdx.y, gridDim.x * gridDim.z, e.err);
      named_bar_sync(1, kConvThreads - 64);
      if (trace && j == 0 && threadIdx.x == 64) trace[19] = clock64();
      bnf_coefficients(blockIdx.x == 0 && blockIdx.z == 0);
      if (trace && j == 0 && threadIdx.x == 64) trace[20] = clock64();
      tc_fence_after();
      if (finalize) {
        for (int ch = ch_begin; ch < ch_end; ++ch) {
          const int col0 = colbase + ch * 16;
          uint32_t r[16];
          tmem_ld16(taddr + ch * 16, r);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          if (valid && col0 < n_valid) {
            if (out_preact != nullptr) store_bf16x16(out_preact, out_base + col0, v);   // raw conv output for backward
            bn_apply_store(v, ch, col0);
          }
        }
      }
      if (trace && j == 0 && threadIdx.x == 64) trace[21] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
    }
    if (bnr_apply) {
      // ---- fused BatchNorm-backward apply: every CTA of this N tile has added sum(dz), sum(dz*x) of its rows
      named_bar_sync(1, kConvThreads - 64);
      if (threadIdx.x == 64) grid_arrive_and_wait(e.bnf_counter + blockIdx.y, gridDim.x * gridDim.z, e.err);
      named_bar_sync(1, kConvThreads - 64);
      float* cA = scratch;                 // [groups][256] each; the reduction scratch is free now
      float* cB = scratch + kMaxBnGroups * 256;
      float* cC = scratch + 2 * kMaxBnGroups * 256;
      {
        const int et = threadIdx.x - 64;
        const bool publish = blockIdx.x == 0 && blockIdx.z == 0;
        if (et < p.block_n) {
          const int c = colbase + et;
          const bool cv = c < e.bnr_c;
          const float gm = cv ? __ldg(e.bnr_gamma + c) : 0.f;
          const float inv_m = 1.f / static_cast<float>(e.bnr_count);
          float dbeta = 0.f, dgamma = 0.f;
          for (int g = 0; g < n_groups; ++g) {
            float A = 0.f, Bx = 0.f, Cc = 0.f;
            if (cv) {
              const float* co = e.bnr_coef + static_cast<long long>(g) * 4 * e.bnr_c;
              const float mu = __ldg(co + 2 * e.bnr_c + c), is = __ldg(co + 3 * e.bnr_c + c);
              const float* sp = e.stats_partial + (static_cast<long long>(g) * e.stats_ld + c) * 2;
              const float s1 = __ldcg(sp);
              const float s2 = is * (__ldcg(sp + 1) - mu * s1);      // sum(dz * xhat) from sum(dz * x)
              const float c2 = s1 * inv_m, c3 = s2 * inv_m;
              A = gm * is;
              Bx = -A * c3 * is;
              Cc = -A * (c2 - mu * is * c3);
              dbeta += s1;
              dgamma += s2;
            }
            cA[g * 256 + et] = A;
            cB[g * 256 + et] = Bx;
            cC[g * 256 + et] = Cc;
          }
          if (publish && cv) {
            if (e.bnr_dbeta != nullptr) e.bnr_dbeta[c] = dbeta;
            if (e.bnr_dgamma != nullptr) e.bnr_dgamma[c] = dgamma;
          }
        }
        if (publish && blockIdx.y == 0 && et == 0 && e.bnr_dalpha != nullptr && e.dalpha_partial != nullptr)
          *e.bnr_dalpha = __ldcg(e.dalpha_partial);
      }
      named_bar_sync(1, kConvThreads - 64);
      tc_fence_after();
      if (finalize) {
        for (int ch = ch_begin; ch < ch_end; ++ch) {
          const int col0 = colbase + ch * 16;
          uint32_t r[16];
          tmem_ld16(taddr + ch * 16, r);
          tmem_ld_wait();
          if (valid && col0 < n_valid) {
            float xr[16], dx[16];
            load_bf16x16(bnr_x, aux_base + col0, xr);
            const float* a = cA + grp * 256 + ch * 16;
            const float* b = cB + grp * 256 + ch * 16;
            const float* c = cC + grp * 256 + ch * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) dx[i] = a[i] * __uint_as_float(r[i]) + b[i] * xr[i] + c[i];
            store_bf16x16(e.bnr_dx, out_base + col0, dx);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
    }
    if (trace && j == 0 && threadIdx.x == 64) trace[6] = clock64();
    if (!PERS) break;
    }  // tile loop
  }

  tc_fence_before();
  __syncthreads();
  // the peer's last multicast commits still arrive on this CTA's barriers: leave together
  if (mc) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (trace && threadIdx.x == 0) trace[7] = clock64();
}

template <int A_MODE, bool FAST = false>
__global__ void __launch_bounds__(kConvThreads, 2) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  conv_body<A_MODE, false, FAST>(p, blockIdx.z, gridDim.z);
}

// Persistent weight-stationary variant (im2col convs only): see the comment above producer_role.
template <bool FAST>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_persistent_kernel(const __grid_constant__ ConvParams p) {
  conv_body<0, true, FAST>(p, 0, 1);
}

// FAST instantiations need 64-channel K chunks (the persistent one also takes 32) and run without trace stamps /
// attribution hooks.
static bool fast_ok(const ConvParams& p) {
  return p.a_mode == 0 && (p.block_k == 64 || (p.block_k == 32 && p.persistent)) && p.epi.trace == nullptr && p.debug == 0;
}

// Up to four independent im2col convs of the same tile grid in ONE launch (blockIdx.z selects the member): the four
// output-parity classes of a stride-2 data gradient, which would otherwise be four latency-bound launches in a row.
template <bool FAST>
__global__ void __launch_bounds__(kConvThreads, 2) conv_igemm_group_kernel(const __grid_constant__ ConvGroup g) {
  const ConvParams& p = g.p[blockIdx.z];
  if (static_cast<int>(blockIdx.x) * kBlockM >= p.M_total || static_cast<int>(blockIdx.y) >= g.tiles_n[blockIdx.z]) return;
  conv_body<0, false, FAST>(p, 0, 1);
}

size_t conv_igemm_smem_bytes(const ConvParams& p) {
  return 1024 + kHeaderBytes + p.b_res_bytes + static_cast<size_t>(p.stages) * p.stage_bytes + p.extra_bytes;
}

// CTAs of this conv's kernel instantiation that the device can hold at once (occupancy x SM count): the fused
// training BatchNorm's grid barrier needs the whole grid resident. Returns -1 on a failed query.
template <typename Kernel>
static int max_coresident(Kernel kernel, size_t smem, int tmem_cols, int* per_sm) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -1;
  int n = 0, dev = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kConvThreads, smem) != cudaSuccess) return -1;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  {
    // The runtime's occupancy calculator answers 1 block per SM for ANY kernel that contains tcgen05.alloc, whatever its
    // block size or shared memory (tools/occ_probe: a 12-register kernel with 32 KB gets 1), while the hardware does
    // place two such CTAs on an SM when it is otherwise empty (same probe: 296 CTAs holding 64 TMEM columns and 96 KB
    // each all meet at a device-wide counter). TSR_OCCUPANCY=own counts the resources here instead (registers per warp
    // in units of 256 over the four register files, shared memory + the per-block reservation against 200 KB, threads,
    // TMEM columns) and so lets 149..296-CTA grids use the grid-barrier epilogues. It is OPT-IN and not safe in general:
    // inside the whole-step CUDA graph, where other branches keep SMs busy, the second CTA of some SMs never becomes
    // resident while the first spins at the barrier (watchdog code 6 at B=64; profiles/r02c_two_ctas_per_sm.md).
    cudaFuncAttributes fa;
    int regs_sm = 0, resv = 0, thr_sm = 0;
    if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess &&
        cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&resv, cudaDevAttrReservedSharedMemoryPerBlock, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev) == cudaSuccess) {
      const int warps = kConvThreads / 32;
      const int regs_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
      const int by_regs = (regs_sm / 4 / regs_warp) * 4 / warps;
      const int by_smem = static_cast<int>((200 * 1024) / (smem + fa.sharedSizeBytes + resv));
      const int by_thr = thr_sm / kConvThreads;
      const int by_tmem = tmem_cols > 0 ? 512 / tmem_cols : 1;
      const int own = std::min(std::min(by_regs, by_tmem), std::min(by_smem, by_thr));
      const char* t = getenv("TSR_OCCUPANCY");
      if (own > n && t && t[0] == 'o') n = own;
    }
  }
  if (per_sm) *per_sm = n;
  if (const char* v = getenv("TSR_CONV_VERBOSE"); v && v[0] == '2') {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kernel);
    int smem_sm = 0, regs_sm = 0, resv = 0, blocks_sm = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&resv, cudaDevAttrReservedSharedMemoryPerBlock, dev);
    cudaDeviceGetAttribute(&blocks_sm, cudaDevAttrMaxBlocksPerMultiprocessor, dev);
    int n2[4] = {0, 0, 0, 0};
    const size_t probe[4] = {32768, 65536, 98304, 112 * 1024};
    for (int i = 0; i < 4; ++i) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n2[i], kernel, kConvThreads, probe[i]);
    size_t avail2 = 0, avail1 = 0;
    cudaOccupancyAvailableDynamicSMemPerBlock(&avail2, kernel, 2, kConvThreads);
    cudaOccupancyAvailableDynamicSMemPerBlock(&avail1, kernel, 1, kConvThreads);
    int n256 = 0, n128 = 0, nflag = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n256, kernel, 256, 32768);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n128, kernel, 128, 32768);
    cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&nflag, kernel, kConvThreads, 32768, cudaOccupancyDisableCachingOverride);
    fprintf(stderr, "[tsr] occupancy probes: dyn smem available for 2 blocks/SM %zu, for 1 %zu; blocks at 256 thr %d, 128 thr %d, "
            "flags %d; maxThreadsPerBlock %d binaryVersion %d\n", avail2, avail1, n256, n128, nflag, fa.maxThreadsPerBlock,
            fa.binaryVersion);
    fprintf(stderr, "[tsr] occupancy: regs %d static smem %zu local %zu maxdyn %d | SM: smem %d regs %d reserved/block %d "
            "blocks %d | blocks/SM at 32K %d 64K %d 96K %d 112K %d, at %zu B: %d\n", fa.numRegs, fa.sharedSizeBytes,
            fa.localSizeBytes, fa.maxDynamicSharedSizeBytes, smem_sm, regs_sm, resv, blocks_sm, n2[0], n2[1], n2[2], n2[3],
            smem, n);
  }
  return n * sms;
}

// One launch path for every instantiation: raises the dynamic shared-memory limit of `kernel` once, then launches.
template <typename Kernel, typename Params>
static cudaError_t launch_conv_kernel(Kernel kernel, bool* attr_set, dim3 grid, size_t smem, cudaStream_t stream, bool pdl,
                                      const Params& params, int cluster_y = 1) {
  if (!*attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    *attr_set = true;
  }
  cudaError_t e = launch_kc(kernel, grid, dim3(kConvThreads), smem, stream, pdl, cluster_y, params);
  return e != cudaSuccess ? e : cudaGetLastError();
}

template <int A_MODE, bool FAST>
static cudaError_t launch_mode(const ConvParams& p, dim3 grid, size_t smem, cudaStream_t stream, bool pdl) {
  static bool attr_set = false;
  return launch_conv_kernel(conv_igemm_kernel<A_MODE, FAST>, &attr_set, grid, smem, stream, pdl, p,
                            (A_MODE == 0 && FAST && p.cluster_n == 2) ? 2 : 1);
}

cudaError_t launch_conv_group(const ConvGroup& g, int n, cudaStream_t stream, bool pdl) {
  int tiles_m = 0, tiles_n = 0;
  size_t smem = 0;
  bool fast = true;
  for (int k = 0; k < n; ++k) {
    tiles_m = max(tiles_m, (g.p[k].M_total + kBlockM - 1) / kBlockM);
    tiles_n = max(tiles_n, g.tiles_n[k]);
    smem = max(smem, conv_igemm_smem_bytes(g.p[k]));
    fast = fast && fast_ok(g.p[k]);
  }
  const dim3 grid(tiles_m, tiles_n, n);
  static bool attr_fast = false, attr_slow = false;
  if (fast) return launch_conv_kernel(conv_igemm_group_kernel<true>, &attr_fast, grid, smem, stream, pdl, g);
  return launch_conv_kernel(conv_igemm_group_kernel<false>, &attr_slow, grid, smem, stream, pdl, g);
}

cudaError_t launch_conv_igemm(const ConvParams& p, int tiles_n, int splits, cudaStream_t stream, bool pdl) {
  const int tiles_m = (p.M_total + kBlockM - 1) / kBlockM;
  dim3 grid(p.persistent ? (p.halo ? p.persistent : min(tiles_m, p.persistent)) : tiles_m, tiles_n, splits);
  const size_t smem = conv_igemm_smem_bytes(p);
  const bool fast = fast_ok(p);
  if (p.persistent) {
    static bool attr_fast = false, attr_slow = false;
    if (fast) return launch_conv_kernel(conv_igemm_persistent_kernel<true>, &attr_fast, grid, smem, stream, pdl, p);
    return launch_conv_kernel(conv_igemm_persistent_kernel<false>, &attr_slow, grid, smem, stream, pdl, p);
  }
  if (p.a_mode == 0)
    return fast ? launch_mode<0, true>(p, grid, smem, stream, pdl) : launch_mode<0, false>(p, grid, smem, stream, pdl);
  if (p.a_mode == 1) return launch_mode<1, false>(p, grid, smem, stream, pdl);
  return launch_mode<2, false>(p, grid, smem, stream, pdl);
}

int conv_igemm_max_coresident(const ConvParams& p, int* per_sm) {
  const size_t smem = conv_igemm_smem_bytes(p);
  const bool fast = fast_ok(p);
  if (p.persistent)
    return fast ? max_coresident(conv_igemm_persistent_kernel<true>, smem, p.tmem_cols, per_sm)
                : max_coresident(conv_igemm_persistent_kernel<false>, smem, p.tmem_cols, per_sm);
  if (p.a_mode != 0) return -1;
  return fast ? max_coresident(conv_igemm_kernel<0, true>, smem, p.tmem_cols, per_sm)
              : max_coresident(conv_igemm_kernel<0, false>, smem, p.tmem_cols, per_sm);
}

}  // namespace tsr
