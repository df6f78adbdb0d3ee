// HBM/L2-bound kernels of the path: layout conversion, BatchNorm (finalize / apply / backward), activation
// backward, loss reductions, weight pack / gradient unpack, classifier head. All use 16-byte vector accesses on
// the NHWC bf16 tensors (8 channels per thread) and warp-shuffle / shared-memory reductions.
// Argument conventions of tsr_elt_desc_t (p[], i[], f[]) are documented at each kernel.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/torchsr_b200.h"
#include "launch.h"
#include "ptx.cuh"

namespace tsr {

namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  uint4 q = *reinterpret_cast<const uint4*>(p);
  unpack_bf16x2(q.x, v[0], v[1]);
  unpack_bf16x2(q.y, v[2], v[3]);
  unpack_bf16x2(q.z, v[4], v[5]);
  unpack_bf16x2(q.w, v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = q;
}
__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
  if (act == TSR_ACT_NONE) return z;
  if (act == TSR_ACT_RELU) return fmaxf(z, 0.f);
  return z > 0.f ? z : z * slope;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------- IM2ROW
// p0 = x fp32 NCHW [B,C,H,W], p1 = E bf16 [B,H,W,Epad]
// i: 0 B, 1 C, 2 H, 3 W, 4 KH, 5 KW, 6 ph, 7 pw, 8 sign, 9 Epad
// E[n,h,w,(kh*KW+kw)*C+c] = x[n,c,h+sign*(kh-ph),w+sign*(kw-pw)] (0 outside), columns >= KH*KW*C are 0.
__global__ void im2row_kernel(const float* __restrict__ x, bf16* __restrict__ E, int B, int C, int H, int W, int KH,
                              int KW, int ph, int pw, int sign, int Epad) {
  pdl_sync();
  const int groups = Epad / 8;
  const long long total = static_cast<long long>(B) * H * W * groups;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    long long pix = idx / groups;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = g * 8 + j;
      float val = 0.f;
      if (col < KH * KW * C) {
        const int c = col % C;
        const int t = col / C;
        const int kw = t % KW, kh = t / KW;
        const int hh = h + sign * (kh - ph), ww = w + sign * (kw - pw);
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(x + ((static_cast<long long>(n) * C + c) * H + hh) * W + ww);
      }
      v[j] = val;
    }
    st8(E + pix * Epad + g * 8, v);
  }
}

// The same mapping with compile-time filter constants (the SRGAN generator's 9x9 3-channel input layer sits at the very
// head of the step's critical path): the per-element divisions of the generic kernel become multiplies.
template <int KH, int KW, int C>
__global__ void __launch_bounds__(256) im2row_ct_kernel(const float* __restrict__ x, bf16* __restrict__ E, int B, int H, int W,
                                                        int ph, int pw, int sign, int Epad) {
  pdl_sync();
  const int groups = Epad / 8;
  const long long total = static_cast<long long>(B) * H * W * groups;
  const long long hw = static_cast<long long>(H) * W;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * 256) {
    const int g = static_cast<int>(idx % groups);
    const long long pix = idx / groups;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const long long n = pix / hw;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = g * 8 + j;
      float val = 0.f;
      if (col < KH * KW * C) {
        const int c = col % C, t = col / C;
        const int kw = t % KW, kh = t / KW;
        const int hh = h + sign * (kh - ph), ww = w + sign * (kw - pw);
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(x + (n * C + c) * hw + static_cast<long long>(hh) * W + ww);
      }
      v[j] = val;
    }
    st8(E + pix * Epad + g * 8, v);
  }
}

// Shared-memory tiled variant for large images (x4 inference: the 512x512 input of the 9x9 layer expands to 134 MB of
// rows): a block stages the (TH+KH-1) x (TW+KW-1) x C input patch of a TH x TW pixel tile once, then every thread
// assembles 16-byte row pieces from shared memory - the row writes are the only HBM traffic left (the per-element
// __ldg gathers of the kernel above ran at a fifth of the write bandwidth).
template <int KH, int KW, int C>
__global__ void __launch_bounds__(256) im2row_tiled_kernel(const float* __restrict__ x, bf16* __restrict__ E, int B, int H,
                                                           int W, int ph, int pw, int sign, int Epad) {
  constexpr int TH = 4, TW = 32, PH = TH + KH - 1, PW = TW + KW - 1;
  __shared__ float patch[C][PH][PW + 1];
  pdl_sync();
  const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
  const long long n_tiles = static_cast<long long>(B) * tiles_h * tiles_w;
  const int lo_h = sign > 0 ? -ph : -(KH - 1 - ph), lo_w = sign > 0 ? -pw : -(KW - 1 - pw);
  const int groups = Epad / 8;     // <= 64: a lane owns column group `lane` and, for wide rows, `lane + 32`
  const long long hw = static_cast<long long>(H) * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int off[8], off2[8];             // patch offsets of the lane's columns relative to the pixel's patch origin, -1 = padding
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    for (int k = 0; k < 2; ++k) {
      const int col = (lane + 32 * k) * 8 + j;
      int o = -1;
      if (col < KH * KW * C) {
        const int c = col % C, tt = col / C;
        const int kw = tt % KW, kh = tt / KW;
        o = (c * PH + sign * (kh - ph) - lo_h) * (PW + 1) + sign * (kw - pw) - lo_w;
      }
      if (k == 0) off[j] = o; else off2[j] = o;
    }
  }
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tw = static_cast<int>(t % tiles_w);
    const int th = static_cast<int>((t / tiles_w) % tiles_h);
    const long long n = t / (static_cast<long long>(tiles_w) * tiles_h);
    const int h0 = th * TH, w0 = tw * TW;
    __syncthreads();   // readers of the previous tile
    for (int i = threadIdx.x; i < C * PH * PW; i += 256) {
      const int pwi = i % PW, phi = (i / PW) % PH, c = i / (PW * PH);
      const int hh = h0 + lo_h + phi, ww = w0 + lo_w + pwi;
      patch[c][phi][pwi] =
          (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + (n * C + c) * hw + static_cast<long long>(hh) * W + ww) : 0.f;
    }
    __syncthreads();
    // lane = column group (loop invariant: the eight (channel, tap) sources of its columns), warp = pixel slot
    for (int pl = warp; pl < TH * TW; pl += 8) {
      const int wl = pl % TW, hl = pl / TW;
      const int h = h0 + hl, w = w0 + wl;
      if (h >= H || w >= W) continue;
      const float* base = &patch[0][hl][wl];
      for (int g = lane; g < groups; g += 32) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (g == lane ? off[j] : off2[j]) >= 0 ? base[g == lane ? off[j] : off2[j]] : 0.f;
        st8(E + ((n * H + h) * W + w) * Epad + g * 8, v);
      }
    }
  }
}

// Fast path for the 3-channel 3x3 first layers of the discriminators and VGG: one thread per pixel gathers its 27 taps
// once and writes the whole 64-byte row; the generic kernel above spends most of its time on per-element index
// arithmetic.
template <int KH, int KW, int C>
__global__ void __launch_bounds__(256) im2row_small_kernel(const float* __restrict__ x, bf16* __restrict__ E, int B, int H,
                                                           int W, int ph, int pw, int sign) {
  static_assert(KH * KW * C <= 32, "row must fit 32 columns");
  pdl_sync();
  const long long total = static_cast<long long>(B) * H * W;
  const long long hw = static_cast<long long>(H) * W;
  for (long long pix = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const long long n = pix / hw;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    // fully unrolled: all KH*KW*C loads are independent and in flight together
#pragma unroll
    for (int kh = 0; kh < KH; ++kh) {
      const int hh = h + sign * (kh - ph);
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        const int ww = w + sign * (kw - pw);
        const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
        for (int c = 0; c < C; ++c)
          v[(kh * KW + kw) * C + c] = in ? __ldg(x + (n * C + c) * hw + static_cast<long long>(hh) * W + ww) : 0.f;
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(E + pix * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 o;
      o.x = pack_bf16x2(v[8 * q], v[8 * q + 1]);
      o.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
      o.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
      o.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
      dst[q] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------- GATHER_OUT
// p0 = T ([B,H,W,Tld] fp32 or bf16), p1 = out fp32 NCHW [B,C,H,W], p2 = bias fp32[C] or null
// i: 0 B, 1 C, 2 H, 3 W, 4 KH, 5 KW, 6 ph, 7 pw, 8 sign, 9 Tld, 10 T_is_bf16
// out[n,c,h,w] = bias[c] + sum_{kh,kw} T[n, h+sign*(kh-ph), w+sign*(kw-pw), (kh*KW+kw)*C+c]
__global__ void gather_out_kernel(const void* __restrict__ T, float* __restrict__ out, const float* __restrict__ bias,
                                  int B, int C, int H, int W, int KH, int KW, int ph, int pw, int sign, int Tld,
                                  int t_bf16) {
  pdl_sync();
  const long long total = static_cast<long long>(B) * C * H * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(idx % W);
    const int h = static_cast<int>((idx / W) % H);
    const int c = static_cast<int>((idx / (static_cast<long long>(W) * H)) % C);
    const int n = static_cast<int>(idx / (static_cast<long long>(W) * H * C));
    float acc = bias ? __ldg(bias + c) : 0.f;
    for (int kh = 0; kh < KH; ++kh) {
      const int hh = h + sign * (kh - ph);
      if (hh < 0 || hh >= H) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int ww = w + sign * (kw - pw);
        if (ww < 0 || ww >= W) continue;
        const long long off = ((static_cast<long long>(n) * H + hh) * W + ww) * Tld + (kh * KW + kw) * C + c;
        acc += t_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(T)[off]) : reinterpret_cast<const float*>(T)[off];
      }
    }
    out[idx] = acc;
  }
}

// Tiled variant for the 3-channel 3x3 case with fp32 rows of 32 columns (gradient w.r.t. the discriminator's input in
// the generator step: on the critical path between D's backward and G's backward): a block stages the rows of a
// (TH + 2) x (TW + 2) pixel patch in shared memory with coalesced 128-byte reads, then every thread sums the nine taps
// of one output element - the kernel above reads nine half-used sectors per output straight from L2.
__global__ void __launch_bounds__(256) gather_out3x3_tiled_kernel(const float* __restrict__ T, float* __restrict__ out,
                                                                  const float* __restrict__ bias, int B, int H, int W,
                                                                  int sign) {
  constexpr int TH = 8, TW = 32, PH = TH + 2, PW = TW + 2, C = 3, LD = 32, SLD = 28;   // 27 used columns, padded to 28
  __shared__ float patch[PH * PW][SLD];
  pdl_sync();
  const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
  const long long n_tiles = static_cast<long long>(B) * tiles_h * tiles_w;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tw = static_cast<int>(t % tiles_w);
    const int th = static_cast<int>((t / tiles_w) % tiles_h);
    const long long n = t / (static_cast<long long>(tiles_w) * tiles_h);
    const int h0 = th * TH - 1, w0 = tw * TW - 1;
    __syncthreads();
    // 8 threads per pixel row fetch its 32 floats as float4 (28 kept)
    for (int i = threadIdx.x; i < PH * PW * 8; i += 256) {
      const int q = i & 7, pp = i >> 3;
      const int ph_ = pp / PW, pw_ = pp - ph_ * PW;
      const int hh = h0 + ph_, ww = w0 + pw_;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W)
        v = __ldg(reinterpret_cast<const float4*>(T + ((n * H + hh) * W + ww) * LD) + q);
      if (q < 7) *reinterpret_cast<float4*>(&patch[pp][4 * q]) = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * TH * TW; i += 256) {
      const int wl = i % TW, hl = (i / TW) % TH, c = i / (TW * TH);
      const int h = th * TH + hl, w = tw * TW + wl;
      if (h >= H || w >= W) continue;
      float acc = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
          acc += patch[(hl + 1 + sign * (kh - 1)) * PW + wl + 1 + sign * (kw - 1)][(kh * 3 + kw) * C + c];
      out[((n * C + c) * H + h) * W + w] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------- layout
// NCHW2NHWC: p0 = x fp32 NCHW, p1 = y bf16 NHWC (pixel stride ld, channel offset off); i: 0 B,1 C,2 H,3 W,4 ld,5 off
// channels are processed in groups of 8 (C%8==0)
__global__ void nchw2nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ y, int B, int C, int H, int W, int ld,
                                 int off) {
  pdl_sync();
  const int groups = C / 8;
  const long long hw = static_cast<long long>(H) * W;
  const long long total = static_cast<long long>(B) * groups * hw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = idx % hw;
    const int g = static_cast<int>((idx / hw) % groups);
    const int n = static_cast<int>(idx / (hw * groups));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(x + (static_cast<long long>(n) * C + g * 8 + j) * hw + p);
    st8(y + (static_cast<long long>(n) * hw + p) * ld + off + g * 8, v);
  }
}
// NHWC2NCHW: p0 = x bf16 NHWC (ld, off), p1 = y fp32 NCHW; i: 0 B,1 C,2 H,3 W,4 ld,5 off, 6 accumulate
__global__ void nhwc2nchw_kernel(const bf16* __restrict__ x, float* __restrict__ y, int B, int C, int H, int W, int ld,
                                 int off, int accumulate) {
  pdl_sync();
  const int groups = C / 8;
  const long long hw = static_cast<long long>(H) * W;
  const long long total = static_cast<long long>(B) * groups * hw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = idx % hw;
    const int g = static_cast<int>((idx / hw) % groups);
    const int n = static_cast<int>(idx / (hw * groups));
    float v[8];
    ld8(x + (static_cast<long long>(n) * hw + p) * ld + off + g * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float* dst = y + (static_cast<long long>(n) * C + g * 8 + j) * hw + p;
      *dst = accumulate ? (*dst + v[j]) : v[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------- BN + activation
// BN_ACT: y = act(norm(x)) * x_scale + res * res_scale, the per-channel (scale, shift) being derived on the fly:
//   mode 0: identity (no normalisation)
//   mode 1: training-mode nn.BatchNorm2d from the conv epilogue's column sums p1 = stats [C][2] (sum, sum of squares):
//           biased variance for the normalisation; block 0 also writes coef_out (scale, shift, mean, invstd) for
//           backward and updates running_mean / running_var (momentum, UNBIASED variance) and num_batches_tracked
//   mode 2: eval-mode BatchNorm from running_mean / running_var
// p0 = x bf16 [M][x_ld], p1 = stats, p2 = y bf16 [M][y_ld], p3 = res bf16 or null, p4 = prelu alpha or null, p5 = gamma,
// p6 = beta, p7 = running_mean, p8 = running_var, p9 = num_batches_tracked (int64), p10 = coef_out fp32 [4][C] or null
// i: 0 M, 1 C, 2 x_ld, 3 y_ld, 4 res_ld, 5 act, 6 x_off, 7 y_off, 8 res_off, 9 mode, 10 count (rows behind each group's
// stats), 11 group_rows (> 0: two statistics groups - rows [0, group_rows) and the rest, the real | fake halves of one
// discriminator pass; stats is then [2][C][2], coef_out [2][4][C], the running statistics take two sequential updates
// and num_batches_tracked += 2, exactly as two separate calls would)
// f: 0 leaky slope, 1 res_scale, 2 x_scale, 3 eps, 4 momentum
struct BnActArgs {
  const bf16* x;
  const float* stats;
  bf16* y;
  const bf16* res;
  const float* alpha;
  const float* gamma;
  const float* beta;
  float* rm;
  float* rv;
  long long* nbt;
  float* coef;
  long long M, count, group_rows;
  int C, x_ld, y_ld, res_ld, act, x_off, y_off, res_off, mode;
  float leaky, res_scale, x_scale, eps, momentum;
};

constexpr int kMaxBnC = 512;   // widest BatchNorm on the path (discriminator 512 channels)

__device__ __forceinline__ void lds8(const float* s, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// The per-channel coefficients are derived ONCE per block into shared memory (one or two channels per thread);
// the row loop then costs two 16-byte loads and one store per 8 channels. Two rows are in flight per thread.
__global__ void __launch_bounds__(256) bn_act_kernel(const BnActArgs a) {
  __shared__ __align__(16) float s_sc[2][kMaxBnC], s_sh[2][kMaxBnC];
  pdl_sync();
  const int C = a.C;
  const int n_groups = (a.group_rows > 0 && a.mode == 1) ? 2 : 1;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float gm = (a.mode != 0 && a.gamma) ? __ldg(a.gamma + c) : 1.f;
    const float bt = (a.mode != 0 && a.beta) ? __ldg(a.beta + c) : 0.f;
    float rm = 0.f, rv = 1.f;
    if (a.mode != 0 && a.rm) {
      rm = a.rm[c];
      rv = a.rv[c];
    }
    for (int g = 0; g < n_groups; ++g) {
      float sc = 1.f, sh = 0.f;
      if (a.mode != 0) {
        float mean = rm, var = rv;
        if (a.mode == 1) {
          const float inv_n = 1.f / static_cast<float>(a.count);
          const float* st = a.stats + static_cast<long long>(g) * 2 * C;
          mean = st[2 * c] * inv_n;
          var = fmaxf(st[2 * c + 1] * inv_n - mean * mean, 0.f);
        }
        const float invstd = rsqrtf(var + a.eps);
        sc = gm * invstd;
        sh = bt - mean * sc;
        if (blockIdx.x == 0) {
          if (a.coef) {
            float* co = a.coef + static_cast<long long>(g) * 4 * C;
            co[0 * C + c] = sc;
            co[1 * C + c] = sh;
            co[2 * C + c] = mean;
            co[3 * C + c] = invstd;
          }
          if (a.mode == 1 && a.rm) {
            const float n = static_cast<float>(a.count);
            const float unbiased = a.count > 1 ? var * n / (n - 1.f) : var;
            rm = (1.f - a.momentum) * rm + a.momentum * mean;
            rv = (1.f - a.momentum) * rv + a.momentum * unbiased;
          }
        }
      }
      s_sc[g][c] = sc;
      s_sh[g][c] = sh;
    }
    if (blockIdx.x == 0 && a.mode == 1 && a.rm) {
      a.rm[c] = rm;
      a.rv[c] = rv;
    }
  }
  if (a.mode == 1 && a.nbt && blockIdx.x == 0 && threadIdx.x == 0) *a.nbt += n_groups;
  __syncthreads();
  const int groups = C >> 3;
  const long long total = a.M * groups;
  const float slope = a.act == TSR_ACT_PRELU ? __ldg(a.alpha) : a.leaky;
  const long long idx0 = blockIdx.x * 256ll + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  const int g = static_cast<int>(idx0 % groups);  // constant per thread: the grid stride is a multiple of `groups`
  const long long split = n_groups == 2 ? a.group_rows : a.M;   // rows >= split use the second coefficient set
  float sc[8], sh[8], sc1[8], sh1[8];
  lds8(s_sc[0] + g * 8, sc);
  lds8(s_sh[0] + g * 8, sh);
  lds8(s_sc[n_groups - 1] + g * 8, sc1);
  lds8(s_sh[n_groups - 1] + g * 8, sh1);
  const bf16* xp = a.x + a.x_off + g * 8;
  const bf16* rp = a.res ? a.res + a.res_off + g * 8 : nullptr;
  bf16* yp = a.y + a.y_off + g * 8;
  // four rows in flight per thread: every load of the iteration is issued before the first dependent instruction
  // (one row per iteration left the kernel latency-bound at ~1.1 TB/s, profiles/r02_ncu_full_bn_act_summary.csv)
  constexpr int R = 4;
  for (long long idx = idx0; idx < total; idx += R * stride) {
    long long m[R];
    bool on[R];
    float v[R][8], r[R][8];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      on[k] = idx + k * stride < total;
      m[k] = (idx + k * stride) / groups;
    }
#pragma unroll
    for (int k = 0; k < R; ++k)
      if (on[k]) ld8(xp + m[k] * a.x_ld, v[k]);
    if (rp) {
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (on[k]) ld8(rp + m[k] * a.res_ld, r[k]);
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (!on[k]) continue;
      const bool hi = m[k] >= split;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = act_fwd(v[k][j] * (hi ? sc1[j] : sc[j]) + (hi ? sh1[j] : sh[j]), a.act, slope) * a.x_scale;
        if (rp) t += r[k][j] * a.res_scale;
        v[k][j] = t;
      }
      st8(yp + m[k] * a.y_ld, v[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- BN backward
// BN_BWD_REDUCE: column sums of dz and dz*xhat over all rows, accumulated with red.global.add into sums [C][2]
// (zeroed by the caller), plus the PReLU slope gradient into a scalar.
// p0 = g bf16 (grad wrt act output) [M][g_ld], p1 = x raw bf16 [M][x_ld], p2 = coef (fwd: scale, shift, mean, invstd)
// or null, p3 = alpha or null, p4 = sums [C][2], p5 = dalpha accumulator (scalar) or null, p6 = optional second
// gradient g2 bf16 added to g (same ld)
// i: 0 M, 1 C, 2 act, 3 rows_per_block, 4 g_ld, 5 x_ld, 6 has_bn; f: 0 leaky, 1 gscale (0 = 1: scales g (+g2) on load)
// dz = g * act'(z), z = x*scale+shift (has_bn) or x; xhat = (x-mean)*invstd. ACT_LEAKY/RELU without BN take
// x = activation output (sign preserved).
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const bf16* __restrict__ g, const bf16* __restrict__ x,
                                                            const float* __restrict__ coef,
                                                            const float* __restrict__ alpha, float* __restrict__ sums,
                                                            float* __restrict__ dalpha_acc,
                                                            const bf16* __restrict__ g2, long long M, int C, int act,
                                                            int rows_per_block, int g_ld, int x_ld, int has_bn,
                                                            float leaky, float gscale) {
  pdl_sync();
  extern __shared__ float sm[];
  __shared__ __align__(16) float s_co[4][kMaxBnC];
  const int groups = C / 8;
  const int lanes = 256 / groups;  // row lanes (groups <= 256)
  const int g_id = threadIdx.x % groups;
  const int r_id = threadIdx.x / groups;
  const float slope = act == TSR_ACT_PRELU ? __ldg(alpha) : (act == TSR_ACT_RELU ? 0.f : leaky);
  for (int c = threadIdx.x; c < C; c += 256) {
    s_co[0][c] = has_bn ? coef[c] : 1.f;
    s_co[1][c] = has_bn ? coef[C + c] : 0.f;
    s_co[2][c] = has_bn ? coef[2 * C + c] : 0.f;
    s_co[3][c] = has_bn ? coef[3 * C + c] : 1.f;
  }
  __syncthreads();
  float sc[8], sh[8], mu[8], is[8];
  lds8(s_co[0] + g_id * 8, sc);
  lds8(s_co[1] + g_id * 8, sh);
  lds8(s_co[2] + g_id * 8, mu);
  lds8(s_co[3] + g_id * 8, is);
  float s1[8], s2[8], da = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long row1 = min(M, row0 + rows_per_block);
  if (r_id < lanes) {
    for (long long m = row0 + r_id; m < row1; m += lanes) {
      float gv[8], xv[8];
      ld8(g + m * g_ld + g_id * 8, gv);
      ld8(x + m * x_ld + g_id * 8, xv);
      if (g2) {
        float t[8];
        ld8(g2 + m * g_ld + g_id * 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) gv[j] += t[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = xv[j] * sc[j] + sh[j];
        float dz = gv[j] * gscale;
        if (act != TSR_ACT_NONE && z <= 0.f) {
          da += dz * z;
          dz *= slope;
        }
        s1[j] += dz;
        s2[j] += dz * (xv[j] - mu[j]) * is[j];
      }
    }
  }
  float* smp = sm;  // [lane][C][2]
  if (r_id < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      smp[(r_id * C + g_id * 8 + j) * 2 + 0] = s1[j];
      smp[(r_id * C + g_id * 8 + j) * 2 + 1] = s2[j];
    }
  }
  float* sda = sm + lanes * C * 2;
  da = warp_sum(da);
  if ((threadIdx.x & 31) == 0) sda[threadIdx.x >> 5] = da;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += smp[l * C * 2 + i];
    atomicAdd(sums + i, a);
  }
  if (dalpha_acc && threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sda[w];
    atomicAdd(dalpha_acc, t);
  }
}

// BN_BWD_APPLY: dx from g (+g2), the raw conv output x, the forward coefficients and the column sums.
// p0 = g bf16, p1 = x raw bf16, p2 = coef fwd or null, p3 = sums [C][2] or null, p4 = alpha, p5 = dx out bf16,
// p6 = optional g2, p7 = gamma or null, p8 = dgamma out [C] or null, p9 = dbeta / bias-gradient out [C] or null,
// p10 = dalpha out (scalar) or null, p11 = dalpha accumulator (scalar) or null
// i: 0 M, 1 C, 2 act, 3 g_ld, 4 x_ld, 5 dx_ld, 6 has_bn; f: 0 leaky, 1 gscale (0 = 1)
// has_bn: dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)); else dx = dz.  Block 0 publishes the parameter
// gradients (dgamma = sum dz*xhat, dbeta = sum dz, dalpha).
struct BnBwdApplyArgs {
  const bf16* g;
  const bf16* x;
  const float* coef;
  const float* sums;
  const float* alpha;
  bf16* dx;
  const bf16* g2;
  const float* gamma;
  float* dgamma;
  float* dbeta;
  float* dalpha;
  const float* dalpha_acc;
  long long M, group_rows;
  int C, act, g_ld, x_ld, dx_ld, has_bn, raw_sums, pre_act;
  float leaky, gscale;
};

// i[7] raw_sums: sums[c][1] holds sum(dz * x) over the RAW conv output x (as accumulated by a data-gradient conv
// epilogue, conv_igemm.cu "fused BatchNorm-backward reduction") instead of sum(dz * xhat);
// i[8] pre_act: g already is dz (the activation derivative was applied by that epilogue).
// i[9] group_rows (> 0, BatchNorm only): two statistics groups as in BN_ACT - coef is [2][4][C], sums [2][C][2], every
// group's rows are normalised with that group's statistics and row count; dgamma / dbeta are the sums over both groups.
// Per channel the whole backward collapses to dx = A*dz + Bx*x + Cc (derived once per block into shared memory).
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdApplyArgs a) {
  __shared__ __align__(16) float s_co[2][5][kMaxBnC];   // per group: sc, sh, A, Bx, Cc
  pdl_sync();
  const int C = a.C;
  const int n_groups = (a.group_rows > 0 && a.has_bn) ? 2 : 1;
  for (int c = threadIdx.x; c < C; c += 256) {
    float dbeta = 0.f, dgamma = 0.f;
    for (int g = 0; g < n_groups; ++g) {
      const long long rows = n_groups == 2 ? (g == 0 ? a.group_rows : a.M - a.group_rows) : a.M;
      const float inv_m = 1.f / static_cast<float>(rows);
      float sc = 1.f, sh = 0.f, mu = 0.f, is = 1.f;
      if (a.has_bn) {
        const float* co = a.coef + static_cast<long long>(g) * 4 * C;
        sc = co[c];
        sh = co[C + c];
        mu = co[2 * C + c];
        is = co[3 * C + c];
      }
      const float* sums = a.sums ? a.sums + static_cast<long long>(g) * 2 * C : nullptr;
      const float s1 = sums ? sums[2 * c] : 0.f;
      float s2 = sums ? sums[2 * c + 1] : 0.f;
      if (a.raw_sums) s2 = is * (s2 - mu * s1);
      const float A = (a.gamma ? __ldg(a.gamma + c) : 1.f) * is;
      const float c2 = s1 * inv_m, c3 = s2 * inv_m;
      s_co[g][0][c] = sc;
      s_co[g][1][c] = sh;
      s_co[g][2][c] = A;
      s_co[g][3][c] = -A * c3 * is;
      s_co[g][4][c] = -A * (c2 - mu * is * c3);
      dbeta += s1;
      dgamma += s2;
    }
    if (blockIdx.x == 0) {
      if (a.dbeta) a.dbeta[c] = dbeta;
      if (a.dgamma) a.dgamma[c] = dgamma;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.dalpha) *a.dalpha = *a.dalpha_acc;
  __syncthreads();
  const int groups = C >> 3;
  const long long total = a.M * groups;
  const float slope = a.act == TSR_ACT_PRELU ? __ldg(a.alpha) : (a.act == TSR_ACT_RELU ? 0.f : a.leaky);
  const bool mask = a.act != TSR_ACT_NONE && !a.pre_act;
  const long long idx0 = blockIdx.x * 256ll + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  const int gi = static_cast<int>(idx0 % groups);
  const long long split = n_groups == 2 ? a.group_rows : a.M;
  const bf16* gp = a.g + gi * 8;
  const bf16* g2p = a.g2 ? a.g2 + gi * 8 : nullptr;
  const bf16* xp = a.x + gi * 8;
  bf16* op = a.dx + gi * 8;
  const bool need_x = a.has_bn || mask;
  // four rows in flight per thread, all global loads of an iteration issued before the math (see bn_act_kernel)
  constexpr int R = 4;
  for (long long idx = idx0; idx < total; idx += R * stride) {
    long long m[R];
    bool on[R];
    float gv[R][8], xv[R][8];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      on[k] = idx + k * stride < total;
      m[k] = (idx + k * stride) / groups;
    }
#pragma unroll
    for (int k = 0; k < R; ++k)
      if (on[k]) ld8(gp + m[k] * a.g_ld, gv[k]);
    if (need_x) {
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (on[k]) ld8(xp + m[k] * a.x_ld, xv[k]);
    }
    if (g2p) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (!on[k]) continue;
        float t[8];
        ld8(g2p + m[k] * a.g_ld, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) gv[k][j] += t[j];
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (!on[k]) continue;
      const float (*cs)[kMaxBnC] = s_co[m[k] >= split ? 1 : 0];
      float sc[8], sh[8], cA[8], cB[8], cC[8];
      lds8(cs[0] + gi * 8, sc);
      lds8(cs[1] + gi * 8, sh);
      lds8(cs[2] + gi * 8, cA);
      lds8(cs[3] + gi * 8, cB);
      lds8(cs[4] + gi * 8, cC);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = gv[k][j] * a.gscale;
        if (mask && xv[k][j] * sc[j] + sh[j] <= 0.f) d *= slope;
        if (a.has_bn) d = cA[j] * d + cB[j] * xv[k][j] + cC[j];
        gv[k][j] = d;
      }
      st8(op + m[k] * a.dx_ld, gv[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- small finalizers
// COLSUM_FINALIZE: p0 = partial [tiles][ld][2], p1 = out[C]; i: 0 tiles, 1 C, 2 ld, 3 which (0/1), 4 accumulate,
// 5 c4 (> 0: the columns are in PixelShuffle-packed order, column r belongs to channel 4*(r % c4) + r / c4)
__global__ void colsum_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int tiles, int C,
                                       int ld, int which, int accumulate, int c4) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int t = 0; t < tiles; ++t) a += partial[(static_cast<long long>(t) * ld + c) * 2 + which];
  const int o = c4 > 0 ? 4 * (c % c4) + c / c4 : c;
  out[o] = accumulate ? out[o] + a : a;
}
// SUM_FINALIZE: p0 = partial[n], p1 = out scalar; i: 0 n, 1 accumulate; f: 0 scale
__global__ void sum_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int accumulate,
                                    float scale) {
  pdl_sync();
  float t = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) t += partial[i];
  t = warp_sum(t) * scale;
  if (threadIdx.x == 0) *out = accumulate ? (*out + t) : t;
}

// ---------------------------------------------------------------------------------------------- pack / unpack
struct PackIdx {
  int co, ci, kh, kw;
  bool valid;
};
// maps packed coordinates (tap slot t, row r, column c) to the OIHW element
__device__ __forceinline__ PackIdx pack_index(const tsr_pack_entry_t& e, int t, int r, int c) {
  PackIdx o;
  o.valid = true;
  const int Cout = e.cout, Cin = e.cin, KH = e.kh, KW = e.kw;
  auto unshuffle = [&](int row) {
    if (!e.shuffle) return row;
    const int c4 = Cout / 4;
    return 4 * (row % c4) + row / c4;
  };
  switch (e.mode) {
    case TSR_PK_FWD:
      o.kh = t / KW; o.kw = t % KW; o.valid = r < Cout && c < Cin; o.co = unshuffle(r); o.ci = c; break;
    case TSR_PK_T:
      o.kh = t / KW; o.kw = t % KW; o.valid = r < Cin && c < Cout; o.ci = r; o.co = unshuffle(c); break;
    case TSR_PK_ROWK:
      o.kh = t; o.valid = r < Cout && c < KW * Cin; o.co = r; o.kw = c / Cin; o.ci = c % Cin; break;
    case TSR_PK_ROWN:
      o.kh = t; o.valid = r < KW * Cout && c < Cin; o.kw = r / Cout; o.co = r % Cout; o.ci = c; break;
    case TSR_PK_ROWN_T:
      o.kh = t; o.valid = r < Cin && c < KW * Cout; o.ci = r; o.kw = c / Cout; o.co = c % Cout; break;
    case TSR_PK_FULLK: {
      o.valid = r < Cout && c < KH * KW * Cin; o.co = r; const int tt = c / Cin; o.ci = c % Cin; o.kh = tt / KW; o.kw = tt % KW;
      break;
    }
    default:
      o.valid = false;
  }
  return o;
}
__device__ __forceinline__ int find_entry(const tsr_pack_entry_t* tab, int n, long long block) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].block_start <= block) lo = mid; else hi = mid - 1;
  }
  return lo;
}
// PACK_W: p0 = device table of tsr_pack_entry_t; i: 0 n_entries; grid = total blocks, 256 threads x 4 elements
__global__ void pack_w_kernel(const tsr_pack_entry_t* __restrict__ tab, int n) {
  __shared__ float tile[32][65];
  pdl_sync();
  const int ei = find_entry(tab, n, blockIdx.x);
  const tsr_pack_entry_t e = tab[ei];
  const long long base = (static_cast<long long>(blockIdx.x) - e.block_start) * 1024;
  const float* src = reinterpret_cast<const float*>(e.src);
  bf16* dst = reinterpret_cast<bf16*>(e.dst);
  if (e.mode == TSR_PK_LINEAR && e.shuffle == 1) {
    // tiled permutation (shuffle == 1 marks the tiled block mapping, see ops.pack_table): one block = one output row
    // x 32 channels x all HW positions. The source run [c0*HW, (c0+32)*HW) of the row is contiguous (coalesced reads),
    // the destination is HW segments of 32 consecutive bf16 (64-byte stores).
    const int HW = e.kh * e.kw, Cc = e.cin, K = Cc * HW;
    const int chunks = (Cc + 31) / 32;
    const long long b = static_cast<long long>(blockIdx.x) - e.block_start;
    const int row = static_cast<int>(b / chunks), c0 = static_cast<int>(b % chunks) * 32;
    if (row >= e.cout) return;
    const int nc = min(32, Cc - c0);
    const float* s0 = src + static_cast<long long>(row) * K + static_cast<long long>(c0) * HW;
    for (int i = threadIdx.x; i < nc * HW; i += 256) tile[i / HW][i % HW] = s0[i];
    __syncthreads();
    bf16* d0 = dst + static_cast<long long>(row) * e.cols_pad + c0;
    for (int i = threadIdx.x; i < nc * HW; i += 256) {
      const int hw = i / nc, c = i % nc;
      d0[static_cast<long long>(hw) * Cc + c] = __float2bfloat16(tile[c][hw]);
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i >= e.count) continue;
    float v = 0.f;
    if (e.mode == TSR_PK_LINEAR) {
      // rows = out features (cout), cols = in features ordered (h, w, c); kh*kw = Hf*Wf, cin = C
      const int K = e.cin * e.kh * e.kw;
      const int row = static_cast<int>(i / e.cols_pad), col = static_cast<int>(i % e.cols_pad);
      if (row < e.cout && col < K) {
        const int c = col % e.cin, hw = col / e.cin;
        v = src[static_cast<long long>(row) * K + static_cast<long long>(c) * e.kh * e.kw + hw];
      }
    } else {
      const int c = static_cast<int>(i % e.cols_pad);
      const long long rr = i / e.cols_pad;
      const int r = static_cast<int>(rr % e.rows_pad);
      const int t = static_cast<int>(rr / e.rows_pad);
      const PackIdx o = pack_index(e, t, r, c);
      if (o.valid) v = src[((static_cast<long long>(o.co) * e.cin + o.ci) * e.kh + o.kh) * e.kw + o.kw];
    }
    dst[i] = __float2bfloat16(v);
  }
}
// UNPACK_G: same table; src = fp32 accumulator laid out [tap][cols_pad][rows_pad] as written by the wgrad kernel
// (rows = the dY columns, contiguous, so that its epilogue can use 16-byte vector reductions), i.e.
// acc[(t * cols_pad + c) * rows_pad + r]; dst = fp32 OIHW gradient (overwritten).
// For TSR_PK_T-style entries r/c are swapped by pack_index; wgrad accumulators always use the forward-like modes
// (FWD, ROWK, ROWN, FULLK). count = rows_pad * taps * cols_pad.
__global__ void unpack_g_kernel(const tsr_pack_entry_t* __restrict__ tab, int n) {
  pdl_sync();
  const int ei = find_entry(tab, n, blockIdx.x);
  const tsr_pack_entry_t e = tab[ei];
  const long long base = (static_cast<long long>(blockIdx.x) - e.block_start) * 1024;
  const float* src = reinterpret_cast<const float*>(e.src);
  float* dst = reinterpret_cast<float*>(e.dst);
  const int taps = (e.mode == TSR_PK_FWD) ? e.kh * e.kw : (e.mode == TSR_PK_FULLK ? 1 : e.kh);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i >= e.count) continue;
    const int r = static_cast<int>(i % e.rows_pad);
    const long long rr = i / e.rows_pad;
    const int c = static_cast<int>(rr % e.cols_pad);
    const int t = static_cast<int>(rr / e.cols_pad);
    (void)taps;
    const PackIdx o = pack_index(e, t, r, c);
    if (o.valid) dst[((static_cast<long long>(o.co) * e.cin + o.ci) * e.kh + o.kh) * e.kw + o.kw] = src[i];
  }
}

// ---------------------------------------------------------------------------------------------- Adam + pack
// ADAM: p0 = device table of tsr_adam_entry_t, p1 = lr (device float, or null -> f[0]), p2 = step (device float: number
// of steps taken so far, incremented by the kernel), p3 = int block counter (zero between launches)
// i: 0 n_entries, 1 total blocks; f: 0 lr (when p1 is null), 1 beta1, 2 beta2, 3 eps, 4 1-beta1, 5 1-beta2
// torch.optim.Adam semantics (torch/optim/_functional / fused kernel): exp_avg = lerp(exp_avg, g, 1-b1);
// exp_avg_sq = b2*exp_avg_sq + (1-b2)*g*g; p -= (lr / (1-b1^t)) * exp_avg / (sqrt(exp_avg_sq)/sqrt(1-b2^t) + eps).
// The same pass writes the bf16 operand copies of the updated weights in the layouts the conv kernels read, so no
// separate pack pass over the parameters is needed after an optimizer step.
constexpr int kAdamTiles = 4;   // LINEAR mode: tiles per block
struct AdamCoef {
  float w1, b2, omb2, step_size, inv_bc2_sqrt, eps;
};
__device__ __forceinline__ float adam_update(float& p, float g, float& m, float& v, const AdamCoef& c) {
  m = m + c.w1 * (g - m);
  v = c.b2 * v + c.omb2 * g * g;
  const float denom = sqrtf(v) * c.inv_bc2_sqrt + c.eps;
  p = p - c.step_size * (m / denom);
  return p;
}
__device__ __forceinline__ int find_adam_entry(const tsr_adam_entry_t* tab, int n, long long block) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].block_start <= block) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__global__ void __launch_bounds__(256) adam_pack_kernel(const tsr_adam_entry_t* __restrict__ tab, int n,
                                                        const float* __restrict__ lr_ptr, float lr_val, float beta1,
                                                        float beta2, float eps, float one_minus_b1, float one_minus_b2,
                                                        float* step_ptr, int* counter) {
  __shared__ float tile[32][65];
  pdl_sync();
  const tsr_adam_entry_t e = tab[find_adam_entry(tab, n, blockIdx.x)];
  const float t = *reinterpret_cast<volatile float*>(step_ptr) + 1.f;
  const float lr = lr_ptr ? *lr_ptr : lr_val;
  AdamCoef c;
  c.w1 = one_minus_b1;      // 1 - beta computed in double on the host (1.f - 0.999f loses 1e-5 relative)
  c.b2 = beta2;
  c.omb2 = one_minus_b2;
  c.step_size = lr / (1.f - powf(beta1, t));
  c.inv_bc2_sqrt = 1.f / sqrtf(1.f - powf(beta2, t));
  c.eps = eps;
  const long long b = static_cast<long long>(blockIdx.x) - e.block_start;
  if (e.mode == TSR_AD_PLAIN) {
    const long long i0 = b * 1024 + threadIdx.x * 4;
    if (i0 + 3 < e.numel && ((reinterpret_cast<uintptr_t>(e.p) | reinterpret_cast<uintptr_t>(e.g)) & 15) == 0) {
      float4 P = *reinterpret_cast<float4*>(e.p + i0), M = *reinterpret_cast<float4*>(e.m + i0),
             V = *reinterpret_cast<float4*>(e.v + i0);
      const float4 G = *reinterpret_cast<const float4*>(e.g + i0);
      adam_update(P.x, G.x, M.x, V.x, c);
      adam_update(P.y, G.y, M.y, V.y, c);
      adam_update(P.z, G.z, M.z, V.z, c);
      adam_update(P.w, G.w, M.w, V.w, c);
      *reinterpret_cast<float4*>(e.p + i0) = P;
      *reinterpret_cast<float4*>(e.m + i0) = M;
      *reinterpret_cast<float4*>(e.v + i0) = V;
    } else {
      for (long long i = i0; i < i0 + 4 && i < static_cast<long long>(e.numel); ++i) adam_update(e.p[i], e.g[i], e.m[i], e.v[i], c);
    }
  } else if (e.mode == TSR_AD_CONV_TILE) {
    // 3x3 conv weight, tile = 16 output channels x 32 input channels x 9 taps: in OIHW that is 16 contiguous runs of
    // 288 floats, so p / g / m / v move as fully coalesced float4 (the per-(co,ci) mapping below touches every 32-byte
    // sector nine times: ncu measured 21 sectors per request and 0.9 TB/s). The updated values are staged in shared
    // memory as bf16 and leave as 64-byte rows of the forward pack ([tap][co][ci]) and 32-byte rows of the transposed
    // pack ([tap][ci][co]).
    __shared__ __align__(16) bf16 sp[16][32 * 9 + 8];
    const int ci_tiles = e.cin >> 5;
    const int co0 = static_cast<int>(b / ci_tiles) * 16, ci0 = static_cast<int>(b % ci_tiles) * 32;
    for (int q = threadIdx.x; q < 16 * 72; q += 256) {
      const int r = q / 72, f = q - r * 72;                  // run (output channel), float4 inside the run
      const long long off = (static_cast<long long>(co0 + r) * e.cin + ci0) * 9 + f * 4;
      float4 P = *reinterpret_cast<float4*>(e.p + off), M = *reinterpret_cast<float4*>(e.m + off),
             V = *reinterpret_cast<float4*>(e.v + off);
      const float4 G = *reinterpret_cast<const float4*>(e.g + off);
      adam_update(P.x, G.x, M.x, V.x, c);
      adam_update(P.y, G.y, M.y, V.y, c);
      adam_update(P.z, G.z, M.z, V.z, c);
      adam_update(P.w, G.w, M.w, V.w, c);
      *reinterpret_cast<float4*>(e.p + off) = P;
      *reinterpret_cast<float4*>(e.m + off) = M;
      *reinterpret_cast<float4*>(e.v + off) = V;
      uint2 h;
      h.x = pack_bf16x2(P.x, P.y);
      h.y = pack_bf16x2(P.z, P.w);
      *reinterpret_cast<uint2*>(&sp[r][f * 4]) = h;           // sp[r][ci_local * 9 + tap]
    }
    __syncthreads();
    bf16* df = reinterpret_cast<bf16*>(e.dst_fwd);
    bf16* dt = reinterpret_cast<bf16*>(e.dst_t);
    if (df) {
      // rows (tap, co): 32 input channels = 64 bytes; one thread = two channels
      for (int i = threadIdx.x; i < 9 * 16 * 16; i += 256) {
        const int pr = i & 15, row = i >> 4;
        const int r = row & 15, k = row >> 4;
        __nv_bfloat162 v2;
        v2.x = sp[r][(2 * pr) * 9 + k];
        v2.y = sp[r][(2 * pr + 1) * 9 + k];
        *reinterpret_cast<__nv_bfloat162*>(df + (static_cast<long long>(k) * e.rows_fwd + co0 + r) * e.cols_fwd + ci0 + 2 * pr) = v2;
      }
    }
    if (dt) {
      // rows (tap, ci): 16 output channels = 32 bytes; one thread = two channels
      for (int i = threadIdx.x; i < 9 * 32 * 8; i += 256) {
        const int pr = i & 7, row = i >> 3;
        const int ci = row & 31, k = row >> 5;
        __nv_bfloat162 v2;
        v2.x = sp[2 * pr][ci * 9 + k];
        v2.y = sp[2 * pr + 1][ci * 9 + k];
        *reinterpret_cast<__nv_bfloat162*>(dt + (static_cast<long long>(k) * e.rows_t + ci0 + ci) * e.cols_t + co0 + 2 * pr) = v2;
      }
    }
  } else if (e.mode == TSR_AD_CONV) {
    // one thread = one (co, ci) pair = kk consecutive OIHW elements; lanes walk ci, so the forward pack
    // ([tap][co'][ci]) gets contiguous 64-byte stores per tap
    const long long pair = b * 256 + threadIdx.x;
    if (pair < static_cast<long long>(e.cout) * e.cin) {
      const int co = static_cast<int>(pair / e.cin), ci = static_cast<int>(pair - static_cast<long long>(co) * e.cin);
      int cop = co;
      if (e.shuffle) {
        const int c4 = e.cout / 4;
        cop = (co & 3) * c4 + (co >> 2);   // inverse of co = 4*(r % c4) + r / c4
      }
      const long long base = pair * e.kk;
      bf16* df = reinterpret_cast<bf16*>(e.dst_fwd);
      bf16* dt = reinterpret_cast<bf16*>(e.dst_t);
      if (e.kk == 9) {
        // 3x3: all 36 loads in flight before the first dependent instruction
        float P[9], G[9], M[9], V[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          P[k] = e.p[base + k];
          G[k] = e.g[base + k];
          M[k] = e.m[base + k];
          V[k] = e.v[base + k];
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          adam_update(P[k], G[k], M[k], V[k], c);
          e.p[base + k] = P[k];
          e.m[base + k] = M[k];
          e.v[base + k] = V[k];
          const bf16 h = __float2bfloat16(P[k]);
          if (df) df[(static_cast<long long>(k) * e.rows_fwd + cop) * e.cols_fwd + ci] = h;
          if (dt) dt[(static_cast<long long>(k) * e.rows_t + ci) * e.cols_t + cop] = h;
        }
      } else {
        for (int k = 0; k < e.kk; ++k) {
          float P = e.p[base + k], M = e.m[base + k], V = e.v[base + k];
          adam_update(P, e.g[base + k], M, V, c);
          e.p[base + k] = P;
          e.m[base + k] = M;
          e.v[base + k] = V;
          const bf16 h = __float2bfloat16(P);
          if (df) df[(static_cast<long long>(k) * e.rows_fwd + cop) * e.cols_fwd + ci] = h;
          if (dt) dt[(static_cast<long long>(k) * e.rows_t + ci) * e.cols_t + cop] = h;
        }
      }
    }
  } else {
    // LINEAR: one tile = one output row x 32 channels x all HW positions (a contiguous run of the parameter); a block
    // walks kAdamTiles consecutive tiles of a row so that the per-block prologue is amortised
    const int HW = e.kk, Cc = e.cin, K = Cc * HW;
    const int chunks = (Cc + 31) / 32;
    const int groups = (chunks + kAdamTiles - 1) / kAdamTiles;
    const int row = static_cast<int>(b / groups);
    const int ch0 = static_cast<int>(b % groups) * kAdamTiles;
    for (int ch = ch0; ch < min(ch0 + kAdamTiles, chunks); ++ch) {
      const int c0 = ch * 32;
      const int nc = min(32, Cc - c0);
      const int cnt = nc * HW;
      const long long off = static_cast<long long>(row) * K + static_cast<long long>(c0) * HW;
      const bool vec = (cnt & 3) == 0 && (off & 3) == 0 &&
                       ((reinterpret_cast<uintptr_t>(e.p) | reinterpret_cast<uintptr_t>(e.g) |
                         reinterpret_cast<uintptr_t>(e.m) | reinterpret_cast<uintptr_t>(e.v)) & 15) == 0;
      if (vec) {
        for (int q = threadIdx.x; q < cnt / 4; q += 256) {
          float4 P = *reinterpret_cast<float4*>(e.p + off + 4 * q), M = *reinterpret_cast<float4*>(e.m + off + 4 * q),
                 V = *reinterpret_cast<float4*>(e.v + off + 4 * q);
          const float4 G = *reinterpret_cast<const float4*>(e.g + off + 4 * q);
          adam_update(P.x, G.x, M.x, V.x, c);
          adam_update(P.y, G.y, M.y, V.y, c);
          adam_update(P.z, G.z, M.z, V.z, c);
          adam_update(P.w, G.w, M.w, V.w, c);
          *reinterpret_cast<float4*>(e.p + off + 4 * q) = P;
          *reinterpret_cast<float4*>(e.m + off + 4 * q) = M;
          *reinterpret_cast<float4*>(e.v + off + 4 * q) = V;
          const float pv[4] = {P.x, P.y, P.z, P.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = 4 * q + j;
            tile[i / HW][i % HW] = pv[j];
          }
        }
      } else {
        for (int i = threadIdx.x; i < cnt; i += 256) {
          float P = e.p[off + i], M = e.m[off + i], V = e.v[off + i];
          adam_update(P, e.g[off + i], M, V, c);
          e.p[off + i] = P;
          e.m[off + i] = M;
          e.v[off + i] = V;
          tile[i / HW][i % HW] = P;
        }
      }
      __syncthreads();
      bf16* d0 = reinterpret_cast<bf16*>(e.dst_fwd) + static_cast<long long>(row) * K + c0;
      if ((nc & 1) == 0) {
        // two channels per thread: 4-byte stores, 64 bytes contiguous per (hw, 32-channel chunk)
        const int half = nc >> 1;
        for (int i = threadIdx.x; i < half * HW; i += 256) {
          const int hw = i / half, cc = (i % half) * 2;
          *reinterpret_cast<uint32_t*>(d0 + static_cast<long long>(hw) * Cc + cc) = pack_bf16x2(tile[cc][hw], tile[cc + 1][hw]);
        }
      } else {
        for (int i = threadIdx.x; i < cnt; i += 256) {
          const int hw = i / nc, cc = i % nc;
          d0[static_cast<long long>(hw) * Cc + cc] = __float2bfloat16(tile[cc][hw]);
        }
      }
      __syncthreads();
    }
  }
  // the block that finishes last advances the step counter (every block read it at its start)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(counter, 1) == static_cast<int>(gridDim.x) - 1) {
      *counter = 0;
      *step_ptr = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------- Linear wgrad
// p0 = dpre fp32 [B][Nf] (gradient wrt pre-activation of the Linear), p1 = X fp32 [B][K] in the parameter's (c,h,w)
// column order (the NHWC2NCHW kernel produces it), p2 = dW fp32 [Nf][K], p3 = db fp32 [Nf]
// i: 0 B, 1 Nf, 2 K
// Block = 256 consecutive columns x 16 output features: every global access is coalesced; HBM-bound on the dW write.
constexpr int kLwF = 16;
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ X,
                                                           float* __restrict__ dW, float* __restrict__ db, int B,
                                                           int Nf, int K) {
  pdl_sync();
  extern __shared__ float sd[];  // [kLwF][B]
  const int n0 = blockIdx.y * kLwF;
  for (int i = threadIdx.x; i < kLwF * B; i += 256) {
    const int f = i / B, b = i % B;
    sd[i] = (n0 + f < Nf) ? dpre[static_cast<long long>(b) * Nf + n0 + f] : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k < K) {
    float acc[kLwF];
#pragma unroll
    for (int f = 0; f < kLwF; ++f) acc[f] = 0.f;
    for (int b = 0; b < B; ++b) {
      const float xv = __ldg(X + static_cast<long long>(b) * K + k);
#pragma unroll
      for (int f = 0; f < kLwF; ++f) acc[f] += sd[f * B + b] * xv;
    }
#pragma unroll
    for (int f = 0; f < kLwF; ++f)
      if (n0 + f < Nf) dW[static_cast<long long>(n0 + f) * K + k] = acc[f];
  }
  if (blockIdx.x == 0 && db && threadIdx.x < kLwF && n0 + threadIdx.x < Nf) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += sd[threadIdx.x * B + b];
    db[n0 + threadIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------------------------- losses
// LOSS: p0 = a fp32, p1 = b fp32, p2 = partial out [blocks], p3 = grad out fp32 (d loss / d a) or null
// i: 0 n, 1 kind (0 = MSE, 1 = L1); f: 0 grad scale (upstream * 1/n)
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                   float* __restrict__ partial, float* __restrict__ grad, long long n,
                                                   int kind, float gscale) {
  pdl_sync();
  __shared__ float sw[8];
  float acc = 0.f;
  for (long long i = (blockIdx.x * 256ll + threadIdx.x) * 4; i < n; i += static_cast<long long>(gridDim.x) * 1024) {
    if (i + 3 < n) {
      const float4 x = *reinterpret_cast<const float4*>(a + i), y = *reinterpret_cast<const float4*>(b + i);
      const float d[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
      float gq[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc += kind == 0 ? d[j] * d[j] : fabsf(d[j]);
        gq[j] = kind == 0 ? 2.f * d[j] * gscale : (d[j] > 0.f ? gscale : (d[j] < 0.f ? -gscale : 0.f));
      }
      if (grad) *reinterpret_cast<float4*>(grad + i) = make_float4(gq[0], gq[1], gq[2], gq[3]);
    } else {
      for (long long j = i; j < n; ++j) {
        const float d = a[j] - b[j];
        acc += kind == 0 ? d * d : fabsf(d);
        if (grad) grad[j] = kind == 0 ? 2.f * d * gscale : (d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f));
      }
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sw[w];
    partial[blockIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------------------------- adversarial losses
// GAN_LOSS: the (relativistic) adversarial criteria of both trainers in ONE single-block launch - forward value and
// the gradient w.r.t. the discriminator outputs (saved for backward), labels are constants (no label tensors):
//   mode 0  nn.BCELoss on probabilities (srgan/trainer.py:164,446-448,456):
//           L = mean_i bce(a_i, ya) [+ mean_j bce(b_j, yb)],  bce(p,y) = -(y*max(log p,-100) + (1-y)*max(log(1-p),-100)),
//           d bce/dp = (p - y) / max((1-p)*p, 1e-12)   (ATen binary_cross_entropy_backward)
//   mode 1  relativistic average GAN, discriminator side (esrgan/trainer.py:451-453):
//           L = mean_i bcewl(a_i - mean(b), ya) + mean_j bcewl(b_j - mean(a), yb); gradients flow through both means
//   mode 2  relativistic average GAN, generator side (esrgan/trainer.py:463-468): L = mean_i bcewl(a_i - mean(b), ya),
//           b is a constant (computed under no_grad)
//   bcewl(x,y) = max(x,0) - x*y + log1p(exp(-|x|)),  d/dx = sigmoid(x) - y
// p0 = a fp32 [na], p1 = b fp32 [nb] or null, p2 = loss out (scalar), p3 = dL/da out [na], p4 = dL/db out [nb] or null,
// p5 = optional per-element targets of a, fp32 [na] (the nn.BCELoss(input, target) form; null -> the constant ya)
// i: 0 na, 1 nb, 2 mode; f: 0 ya, 1 yb, 2 scale (multiplies the loss and both gradients)
__device__ __forceinline__ float block_sum_256(float v, float* sw) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += sw[w];
  return t;
}
__device__ __forceinline__ float bce_prob(float p, float y) {
  return -(y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(logf(1.f - p), -100.f));
}
__device__ __forceinline__ float bce_prob_grad(float p, float y) { return (p - y) / fmaxf((1.f - p) * p, 1e-12f); }
__device__ __forceinline__ float bce_logit(float x, float y) { return fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float bce_logit_grad(float x, float y) { return 1.f / (1.f + expf(-x)) - y; }

__global__ void __launch_bounds__(256) gan_loss_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       float* __restrict__ loss, float* __restrict__ ga,
                                                       float* __restrict__ gb, const float* __restrict__ ta,
                                                       int na, int nb, int mode, float ya, float yb, float scale) {
  pdl_sync();
  __shared__ float sw[8];
  const int t = threadIdx.x;
  float ma = 0.f, mb = 0.f;
  if (mode != 0) {
    float sa = 0.f, sb = 0.f;
    for (int i = t; i < na; i += 256) sa += a[i];
    for (int j = t; j < nb; j += 256) sb += b[j];
    ma = block_sum_256(sa, sw) / static_cast<float>(max(na, 1));
    mb = block_sum_256(sb, sw) / static_cast<float>(max(nb, 1));
  }
  const bool b_term = b != nullptr && mode != 2;
  const float inv_a = 1.f / static_cast<float>(max(na, 1)), inv_b = 1.f / static_cast<float>(max(nb, 1));
  float la = 0.f, lb = 0.f, da = 0.f, db = 0.f;   // loss sums, sums of the per-element derivatives
  for (int i = t; i < na; i += 256) {
    const float x = a[i] - (mode != 0 ? mb : 0.f);
    const float y = ta ? ta[i] : ya;
    la += mode == 0 ? bce_prob(x, y) : bce_logit(x, y);
    da += mode == 0 ? bce_prob_grad(x, y) : bce_logit_grad(x, y);
  }
  if (b_term) {
    for (int j = t; j < nb; j += 256) {
      const float x = b[j] - (mode != 0 ? ma : 0.f);
      lb += mode == 0 ? bce_prob(x, yb) : bce_logit(x, yb);
      db += mode == 0 ? bce_prob_grad(x, yb) : bce_logit_grad(x, yb);
    }
  }
  la = block_sum_256(la, sw);
  lb = block_sum_256(lb, sw);
  da = block_sum_256(da, sw);
  db = block_sum_256(db, sw);
  if (t == 0) *loss = scale * (la * inv_a + (b_term ? lb * inv_b : 0.f));
  // mode 1: a_i also enters every b-term through mean(a) (and vice versa): - (1/na) * mean_j bcewl'(b_j - mean a)
  const float cross_a = (mode == 1 && b_term) ? db * inv_b * inv_a : 0.f;
  const float cross_b = mode == 1 ? da * inv_a * inv_b : 0.f;
  for (int i = t; i < na; i += 256) {
    const float x = a[i] - (mode != 0 ? mb : 0.f);
    const float y = ta ? ta[i] : ya;
    const float d = mode == 0 ? bce_prob_grad(x, y) : bce_logit_grad(x, y);
    ga[i] = scale * (d * inv_a - cross_a);
  }
  if (gb != nullptr && b_term) {
    for (int j = t; j < nb; j += 256) {
      const float x = b[j] - (mode != 0 ? ma : 0.f);
      const float d = mode == 0 ? bce_prob_grad(x, yb) : bce_logit_grad(x, yb);
      gb[j] = scale * (d * inv_b - cross_b);
    }
  }
}

// AXPBY_F32: out = a * s * x + b * y on fp32 vectors; s = *p3 when given (a device scalar: the upstream gradient of a
// scalar loss, so that loss backward passes need no host round trip and no ATen multiply), else 1.
// p0 = x, p1 = y or null, p2 = out (may alias x or y), p3 = device scalar or null; i: 0 n; f: 0 a, 1 b
__global__ void __launch_bounds__(256) axpby_f32_kernel(const float* x, const float* y, float* out,
                                                        const float* __restrict__ s, long long n, float a, float b) {
  pdl_sync();
  const float k = a * (s ? __ldg(s) : 1.f);
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                     (y ? reinterpret_cast<uintptr_t>(y) : 0)) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x *= k; v.y *= k; v.z *= k; v.w *= k;
    if (y) {
      const float4 w = reinterpret_cast<const float4*>(y)[i];
      v.x += b * w.x; v.y += b * w.y; v.z += b * w.z; v.w += b * w.w;
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
  for (long long i = n4 * 4 + blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256)
    out[i] = k * x[i] + (y ? b * y[i] : 0.f);
}

// ---------------------------------------------------------------------------------------------- upsample (ESRGAN)
// UPSAMPLE2X: p0 = x bf16 [B,H,W,ld_in], p1 = y bf16 [B,2H,2W,ld_out]; i: 0 B,1 H,2 W,3 C,4 ld_in,5 ld_out
__global__ void upsample2x_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C,
                                  int ld_in, int ld_out) {
  pdl_sync();
  const int groups = C / 8;
  const long long total = static_cast<long long>(B) * 2 * H * 2 * W * groups;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    long long p = idx / groups;
    const int w = static_cast<int>(p % (2 * W));
    const int h = static_cast<int>((p / (2 * W)) % (2 * H));
    const int n = static_cast<int>(p / (4ll * W * H));
    const uint4 q = *reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * H + h / 2) * W + w / 2) * ld_in + g * 8);
    *reinterpret_cast<uint4*>(y + p * ld_out + g * 8) = q;
  }
}
// UPSAMPLE2X_BWD: p0 = dy bf16 [B,2H,2W,ld_in], p1 = dx bf16 [B,H,W,ld_out]; i as above (H,W = coarse dims)
__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int B, int H, int W, int C,
                                      int ld_in, int ld_out) {
  pdl_sync();
  const int groups = C / 8;
  const long long total = static_cast<long long>(B) * H * W * groups;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    long long p = idx / groups;
    const int w = static_cast<int>(p % W);
    const int h = static_cast<int>((p / W) % H);
    const int n = static_cast<int>(p / (static_cast<long long>(W) * H));
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[8];
        ld8(dy + ((static_cast<long long>(n) * 2 * H + 2 * h + i) * 2 * W + 2 * w + j) * ld_in + g * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k];
      }
    st8(dx + p * ld_out + g * 8, acc);
  }
}

// ---------------------------------------------------------------------------------------------- classifier head
// HEAD: p0 = pre1 fp32 [N1][B] (transposed split-K GEMM output, no bias), p1 = b1 [N1], p2 = w2 fp32 [N1], p3 = b2 [1],
// p4 = out fp32 [B], p5 = h1 out fp32 [B][N1] (post-LeakyReLU, saved for backward), p6 = h1 bf16 out or null
// i: 0 B, 1 N1, 2 sigmoid; f: 0 leaky
// discriminator.py:64-69: Linear -> LeakyReLU(0.2) -> Linear(N1,1) -> Sigmoid (SRGAN) / logits (ESRGAN)
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ pre1, const float* __restrict__ b1,
                                                   const float* __restrict__ w2, const float* __restrict__ b2,
                                                   float* __restrict__ out, float* __restrict__ h1, int B, int N1,
                                                   int sigmoid, float leaky) {
  pdl_sync();
  __shared__ float sw[8];
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int k = threadIdx.x; k < N1; k += 256) {
    float z = pre1[static_cast<long long>(k) * B + b] + b1[k];
    z = z > 0.f ? z : z * leaky;
    h1[static_cast<long long>(b) * N1 + k] = z;
    acc += z * w2[k];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = b2[0];
    for (int w = 0; w < 8; ++w) t += sw[w];
    out[b] = sigmoid ? 1.f / (1.f + __expf(-t)) : t;
  }
}
// HEAD_BWD: p0 = gout fp32 [B] (grad wrt head output), p1 = out fp32 [B] (saved output), p2 = h1 fp32 [B][N1],
// p3 = w2 [N1], p4 = dpre1 fp32 [B][N1] out, p5 = dpre1 bf16 [Bpad][N1] out (rows >= B zero-filled by caller),
// p6 = dw2 [N1] out, p7 = db2 [1] out, p8 = db1 [N1] out or null (bias gradient of the first Linear: column sums of dpre1)
// i: 0 B, 1 N1, 2 sigmoid, 3 row stride of dpre1_bf in elements (0 = N1); f: 0 leaky.
// One block per 256 features; loops over batch.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out,
                                                       const float* __restrict__ h1, const float* __restrict__ w2,
                                                       float* __restrict__ dpre1, bf16* __restrict__ dpre1_bf,
                                                       float* __restrict__ dw2, float* __restrict__ db2,
                                                       float* __restrict__ db1, int B, int N1, int sigmoid, int ld_bf,
                                                       float leaky) {
  pdl_sync();
  const int k = blockIdx.x * 256 + threadIdx.x;
  float dw = 0.f, dbs = 0.f, d1 = 0.f;
  for (int b = 0; b < B; ++b) {
    float dl = gout[b];
    if (sigmoid) {
      const float pr = out[b];
      dl *= pr * (1.f - pr);
    }
    dbs += dl;
    if (k < N1) {
      const float h = h1[static_cast<long long>(b) * N1 + k];
      dw += dl * h;
      float d = dl * w2[k];
      if (h <= 0.f) d *= leaky;
      dpre1[static_cast<long long>(b) * N1 + k] = d;
      if (dpre1_bf) dpre1_bf[static_cast<long long>(b) * ld_bf + k] = __float2bfloat16(d);
      d1 += d;
    }
  }
  if (k < N1) dw2[k] = dw;
  if (k < N1 && db1) db1[k] = d1;
  if (k == 0) db2[0] = dbs;
}

// FEAT_T: the flattened feature map in the Linear weight's column order, transposed and cut into 64-row batch chunks -
// the K-major B operand of the Linear weight-gradient GEMM  dW[n][k] = sum_b dpre[b][n] * X[b][k]  (tcgen05, conv_igemm.cu):
//   XT[chunk][k][j] = X[chunk*64 + j][k],  k = c*HW + hw (torch.flatten of NCHW, discriminator.py:86), zero for b >= B.
// A chunk is one K block of the GEMM; an all-gather of every rank's chunks simply appends K blocks (dist.py).
// p0 = act bf16 NHWC [B][HW][ld] (+off), p1 = XT bf16 [chunks][C*HW][64]; i: 0 B, 1 C, 2 HW, 3 ld, 4 off
__global__ void __launch_bounds__(256) feat_t_kernel(const bf16* __restrict__ act, bf16* __restrict__ xt, int B, int C,
                                                     int HW, int ld, int off) {
  pdl_sync();
  const int chunks = (B + 63) / 64;
  const long long total = static_cast<long long>(chunks) * C * HW * 64;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * 256) {
    const int j = static_cast<int>(idx & 63);
    const long long r = idx >> 6;
    const int k = static_cast<int>(r % (static_cast<long long>(C) * HW));
    const int chunk = static_cast<int>(r / (static_cast<long long>(C) * HW));
    const int b = chunk * 64 + j;
    const int c = k / HW, hw = k - c * HW;
    xt[idx] = b < B ? act[(static_cast<long long>(b) * HW + hw) * ld + off + c] : __float2bfloat16(0.f);
  }
}

// ---------------------------------------------------------------------------------------------- input pipeline
// CROP_LR: the reference's per-sample training transform (torchsr/dataset.py:86-99,118-121) for a whole batch in one
// launch, on pre-decoded uint8 images resident in HBM: RandomCrop + horizontal / vertical flip -> HR crop (ToTensor:
// u8 / 255), then ToPILImage -> Resize(crop/4, BICUBIC) -> ToTensor for the LR input. The resize is Pillow's
// ImagingResample for 8-bit images restated bit for bit (src/libImaging/Resample.c: separable, antialiased - support
// 2 * scale, 17 taps at /4 -, integer coefficients with 22 fractional bits, horizontal pass first, every pass rounded
// and clamped to uint8): kk / bounds are the coefficient tables of precompute_coeffs + normalize_coeffs_8bpc, built on
// the host (gpu_data.pil_bicubic_tables).
// p0 = pool u8 (images HWC, RGB), p1 = params int64 [B][6] (byte offset of the image, image width in pixels, x0, y0,
// flip_h, flip_v), p2 = kk int32 [out][ksize], p3 = bounds int32 [out][2] (first tap, tap count), p4 = hr fp32
// [B][3][crop][crop], p5 = lr fp32 [B][3][out][out]; i: 0 B, 1 crop (<= 128), 2 out, 3 ksize. One block per crop.
__global__ void __launch_bounds__(256) crop_lr_kernel(const uint8_t* __restrict__ pool, const long long* __restrict__ params,
                                                      const int* __restrict__ kk, const int* __restrict__ bounds,
                                                      float* __restrict__ hr, float* __restrict__ lr, int crop, int out,
                                                      int ksize) {
  pdl_sync();
  __shared__ uint8_t tmp[128 * 32 * 3];     // horizontally resampled crop [crop][out][3]
  const long long* pr = params + 6ll * blockIdx.x;
  const uint8_t* img = pool + pr[0];
  const int W = static_cast<int>(pr[1]), x0 = static_cast<int>(pr[2]), y0 = static_cast<int>(pr[3]);
  const bool fh = pr[4] != 0, fv = pr[5] != 0;
  auto src = [&](int y, int x) {   // first byte of crop pixel (y, x) in the pool
    const int sy = y0 + (fv ? crop - 1 - y : y), sx = x0 + (fh ? crop - 1 - x : x);
    return img + (static_cast<long long>(sy) * W + sx) * 3;
  };
  float* hr_b = hr + static_cast<long long>(blockIdx.x) * 3 * crop * crop;
  for (int i = threadIdx.x; i < crop * crop; i += 256) {
    const int y = i / crop, x = i - y * crop;
    const uint8_t* s = src(y, x);
#pragma unroll
    for (int c = 0; c < 3; ++c) hr_b[(c * crop + y) * crop + x] = static_cast<float>(s[c]) / 255.f;
  }
  // horizontal pass: (y, xx) -> 3 channels
  for (int i = threadIdx.x; i < crop * out; i += 256) {
    const int y = i / out, xx = i - y * out;
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int* k = kk + xx * ksize;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    for (int t = 0; t < n; ++t) {
      const uint8_t* s = src(y, xmin + t);
      const int w = k[t];
      a0 += s[0] * w;
      a1 += s[1] * w;
      a2 += s[2] * w;
    }
    uint8_t* d = tmp + (y * out + xx) * 3;
    d[0] = static_cast<uint8_t>(min(max(a0 >> 22, 0), 255));
    d[1] = static_cast<uint8_t>(min(max(a1 >> 22, 0), 255));
    d[2] = static_cast<uint8_t>(min(max(a2 >> 22, 0), 255));
  }
  __syncthreads();
  float* lr_b = lr + static_cast<long long>(blockIdx.x) * 3 * out * out;
  for (int i = threadIdx.x; i < out * out * 3; i += 256) {
    const int c = i % 3, p = i / 3;
    const int yy = p / out, xx = p - yy * out;
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int* k = kk + yy * ksize;
    int a = 1 << 21;
    for (int t = 0; t < n; ++t) a += tmp[((ymin + t) * out + xx) * 3 + c] * k[t];
    lr_b[(c * out + yy) * out + xx] = static_cast<float>(min(max(a >> 22, 0), 255)) / 255.f;
  }
}

// ---------------------------------------------------------------------------------------------- misc
// AXPBY (bf16): p0 = x, p1 = y or null, p2 = out; i: 0 n (multiple of 8); f: 0 a, 1 b;  out = a*x + b*y
__global__ void axpby_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, bf16* __restrict__ out,
                             long long n, float a, float b) {
  pdl_sync();
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 8; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 8) {
    float v[8];
    ld8(x + i, v);
    if (y) {
      float t[8];
      ld8(y + i, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = a * v[j] + b * t[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= a;
    }
    st8(out + i, v);
  }
}
// MAXPOOL2 (VGG): p0 = x bf16 [B,H,W,C], p1 = y bf16 [B,H/2,W/2,C]; i: 0 B,1 H,2 W,3 C
__global__ void maxpool2_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C) {
  pdl_sync();
  const int groups = C / 8, Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * groups;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    long long p = idx / groups;
    const int w = static_cast<int>(p % Wo);
    const int h = static_cast<int>((p / Wo) % Ho);
    const int n = static_cast<int>(p / (static_cast<long long>(Wo) * Ho));
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[8];
        ld8(x + ((static_cast<long long>(n) * H + 2 * h + i) * W + 2 * w + j) * C + g * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
      }
    st8(y + p * C + g * 8, m);
  }
}
// MAXPOOL2_BWD: p0 = x (pool input) bf16, p1 = y (pool output) bf16, p2 = dy bf16, p3 = dx out bf16; i as MAXPOOL2,
// i4 relu: x is a ReLU output whose own backward is folded in (gradient only where x > 0).
// Gradient goes to the first element equal to the max in (0,0),(0,1),(1,0),(1,1) order (ATen's tie rule).
__global__ void maxpool2_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, const bf16* __restrict__ dy,
                                    bf16* __restrict__ dx, int B, int H, int W, int C, int relu) {
  pdl_sync();
  const int groups = C / 8, Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * groups;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    long long p = idx / groups;
    const int w = static_cast<int>(p % Wo);
    const int h = static_cast<int>((p / Wo) % Ho);
    const int n = static_cast<int>(p / (static_cast<long long>(Wo) * Ho));
    float m[8], d[8];
    bool done[8];
    ld8(y + p * C + g * 8, m);
    ld8(dy + p * C + g * 8, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) done[k] = false;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[8], o[8];
        const long long off = ((static_cast<long long>(n) * H + 2 * h + i) * W + 2 * w + j) * C + g * 8;
        ld8(x + off, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const bool hit = !done[k] && v[k] == m[k];
          o[k] = (hit && !(relu && v[k] <= 0.f)) ? d[k] : 0.f;
          done[k] = done[k] || hit;
        }
        st8(dx + off, o);
      }
  }
}
// ZERO: p0 = dst (16-byte aligned); i: 0 bytes (multiple of 16)
// PACK_GATHER: dst[i] = idx[i] >= 0 ? bf16(src[idx[i]]) : 0 (table-driven operand layouts of single plans)
__global__ void pack_gather_kernel(const float* __restrict__ src, const int* __restrict__ idx, bf16* __restrict__ dst,
                                   long long n) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = __ldg(idx + i);
    dst[i] = __float2bfloat16(k >= 0 ? __ldg(src + k) : 0.f);
  }
}

__global__ void zero_kernel(uint4* __restrict__ dst, long long n16) {
  pdl_sync();
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = z;
}
// CAST: p0 = src, p1 = dst; i: 0 n, 1 dir (0: fp32 -> bf16, 1: bf16 -> fp32)
__global__ void cast_kernel(const void* __restrict__ src, void* __restrict__ dst, long long n, int dir) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (dir == 0)
      reinterpret_cast<bf16*>(dst)[i] = __float2bfloat16(reinterpret_cast<const float*>(src)[i]);
    else
      reinterpret_cast<float*>(dst)[i] = __bfloat162float(reinterpret_cast<const bf16*>(src)[i]);
  }
}

// CHANSUM_NCHW: p0 = x fp32 [B][C][HW], p1 = partial out [splits][C][2] (slot 0 = sum); i: 0 B, 1 C, 2 HW, 3 splits
// grid (C, splits): split s of channel c covers a contiguous slice of the B*HW elements of that channel.
__global__ void __launch_bounds__(256) chansum_nchw_kernel(const float* __restrict__ x, float* __restrict__ partial, int B,
                                                           int C, long long HW, int splits) {
  pdl_sync();
  __shared__ float sw[8];
  const int c = blockIdx.x, s = blockIdx.y;
  const long long total = static_cast<long long>(B) * HW;
  const long long per = (total + splits - 1) / splits;
  const long long lo = s * per, hi = min(total, lo + per);
  float acc = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) {
    const long long b = i / HW, r = i - b * HW;
    acc += __ldg(x + (b * C + c) * HW + r);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sw[w];
    partial[(static_cast<long long>(s) * C + c) * 2 + 0] = t;
    partial[(static_cast<long long>(s) * C + c) * 2 + 1] = 0.f;
  }
}

inline int grid_for(long long work_items, int block = 256, int max_blocks = 148 * 16) {
  long long b = (work_items + block - 1) / block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}

}  // namespace

cudaError_t launch_elt(const tsr_elt_desc_t& d, cudaStream_t st, bool pdl) {
  cudaError_t ce = cudaSuccess;
  const int64_t* i = d.i;
  void* const* p = d.p;
  switch (d.kind) {
    case TSR_E_IM2ROW:
      if (i[9] == 32 && i[1] == 3 && i[4] == 3 && i[5] == 3) {      // 3-channel 3x3 first layers (D, VGG)
        ce = launch_k(im2row_small_kernel<3, 3, 3>, dim3(grid_for(i[0] * i[2] * i[3])), dim3(256), 0, st, pdl,
                      (const float*)p[0], (bf16*)p[1], i[0], i[2], i[3], i[6], i[7], i[8]);
        break;
      }
      if (i[9] == 32 && i[1] == 3 && i[4] == 1 && i[5] == 9) {      // row expansion of dOut for the 9x9 Cout=3 conv
        ce = launch_k(im2row_small_kernel<1, 9, 3>, dim3(grid_for(i[0] * i[2] * i[3])), dim3(256), 0, st, pdl,
                      (const float*)p[0], (bf16*)p[1], i[0], i[2], i[3], i[6], i[7], i[8]);
        break;
      }
      if (i[1] == 3 && i[4] == 9 && i[5] == 9 && i[0] * i[2] * i[3] >= 65536 && i[9] <= 512) {   // the same layer on large images
        const long long n_tiles = i[0] * ((i[2] + 3) / 4) * ((i[3] + 31) / 32);
        ce = launch_k(im2row_tiled_kernel<9, 9, 3>, dim3(static_cast<unsigned>(n_tiles < 148 * 16 ? n_tiles : 148 * 16)), dim3(256),
                      0, st, pdl, (const float*)p[0], (bf16*)p[1], i[0], i[2], i[3], i[6], i[7], i[8], i[9]);
        break;
      }
      if (i[1] == 3 && i[4] == 9 && i[5] == 9) {                      // 9x9 3-channel input layer of the SRGAN generator
        ce = launch_k(im2row_ct_kernel<9, 9, 3>, dim3(grid_for(i[0] * i[2] * i[3] * (i[9] / 8))), dim3(256), 0, st, pdl,
                      (const float*)p[0], (bf16*)p[1], i[0], i[2], i[3], i[6], i[7], i[8], i[9]);
        break;
      }
      ce = launch_k(im2row_kernel, dim3(grid_for(i[0] * i[2] * i[3] * (i[9] / 8))), dim3(256), 0, st, pdl, 
          (const float*)p[0], (bf16*)p[1], i[0], i[1], i[2], i[3], i[4], i[5], i[6], i[7], i[8], i[9]);
      break;
    case TSR_E_GATHER_OUT:
      if (i[1] == 3 && i[4] == 3 && i[5] == 3 && i[6] == 1 && i[7] == 1 && i[9] == 32 && i[10] == 0 &&
          (reinterpret_cast<uintptr_t>(p[0]) & 15) == 0) {
        const long long n_tiles = i[0] * ((i[2] + 7) / 8) * ((i[3] + 31) / 32);
        ce = launch_k(gather_out3x3_tiled_kernel, dim3(static_cast<unsigned>(n_tiles < 148 * 8 ? n_tiles : 148 * 8)), dim3(256), 0,
                      st, pdl, (const float*)p[0], (float*)p[1], (const float*)p[2], i[0], i[2], i[3], i[8]);
        break;
      }
      ce = launch_k(gather_out_kernel, dim3(grid_for(i[0] * i[1] * i[2] * i[3])), dim3(256), 0, st, pdl, p[0], (float*)p[1], (const float*)p[2], i[0],
                                                                          i[1], i[2], i[3], i[4], i[5], i[6], i[7], i[8],
                                                                          i[9], i[10]);
      break;
    case TSR_E_NCHW2NHWC:
      ce = launch_k(nchw2nhwc_kernel, dim3(grid_for(i[0] * (i[1] / 8) * i[2] * i[3])), dim3(256), 0, st, pdl, (const float*)p[0], (bf16*)p[1], i[0],
                                                                               i[1], i[2], i[3], i[4], i[5]);
      break;
    case TSR_E_NHWC2NCHW:
      ce = launch_k(nhwc2nchw_kernel, dim3(grid_for(i[0] * (i[1] / 8) * i[2] * i[3])), dim3(256), 0, st, pdl, (const bf16*)p[0], (float*)p[1], i[0],
                                                                               i[1], i[2], i[3], i[4], i[5], i[6]);
      break;
    case TSR_E_BN_ACT: {
      BnActArgs a;
      a.x = (const bf16*)p[0]; a.stats = (const float*)p[1]; a.y = (bf16*)p[2]; a.res = (const bf16*)p[3];
      a.alpha = (const float*)p[4]; a.gamma = (const float*)p[5]; a.beta = (const float*)p[6]; a.rm = (float*)p[7];
      a.rv = (float*)p[8]; a.nbt = (long long*)p[9]; a.coef = (float*)p[10];
      a.M = i[0]; a.C = i[1]; a.x_ld = i[2]; a.y_ld = i[3]; a.res_ld = i[4]; a.act = i[5]; a.x_off = i[6];
      a.y_off = i[7]; a.res_off = i[8]; a.mode = i[9]; a.count = i[10]; a.group_rows = i[11];
      a.leaky = d.f[0]; a.res_scale = d.f[1]; a.x_scale = d.f[2]; a.eps = d.f[3]; a.momentum = d.f[4];
      if (a.C > kMaxBnC) return cudaErrorInvalidValue;
      ce = launch_k(bn_act_kernel, dim3(grid_for((i[0] * (i[1] / 8) + 3) / 4, 256, 148 * 8)), dim3(256), 0, st, pdl, a);
      break;
    }
    case TSR_E_BN_BWD_REDUCE: {
      const int C = i[1];
      if (C > kMaxBnC) return cudaErrorInvalidValue;
      const int lanes = 256 / (C / 8);
      const long long blocks = (i[0] + i[3] - 1) / i[3];
      const size_t sm = (static_cast<size_t>(lanes) * C * 2 + 8) * sizeof(float);
      ce = launch_k(bn_bwd_reduce_kernel, dim3(blocks), dim3(256), sm, st, pdl, (const bf16*)p[0], (const bf16*)p[1], (const float*)p[2],
                                                    (const float*)p[3], (float*)p[4], (float*)p[5], (const bf16*)p[6],
                                                    i[0], C, i[2], i[3], i[4], i[5], i[6], d.f[0],
                                                    d.f[1] != 0.f ? d.f[1] : 1.f);
      break;
    }
    case TSR_E_BN_BWD_APPLY: {
      BnBwdApplyArgs a;
      a.g = (const bf16*)p[0]; a.x = (const bf16*)p[1]; a.coef = (const float*)p[2]; a.sums = (const float*)p[3];
      a.alpha = (const float*)p[4]; a.dx = (bf16*)p[5]; a.g2 = (const bf16*)p[6]; a.gamma = (const float*)p[7];
      a.dgamma = (float*)p[8]; a.dbeta = (float*)p[9]; a.dalpha = (float*)p[10]; a.dalpha_acc = (const float*)p[11];
      a.M = i[0]; a.C = i[1]; a.act = i[2]; a.g_ld = i[3]; a.x_ld = i[4]; a.dx_ld = i[5]; a.has_bn = i[6];
      a.raw_sums = i[7]; a.pre_act = i[8]; a.group_rows = i[9];
      if (a.C > kMaxBnC) return cudaErrorInvalidValue;
      a.leaky = d.f[0];
      a.gscale = d.f[1] != 0.f ? d.f[1] : 1.f;
      ce = launch_k(bn_bwd_apply_kernel, dim3(grid_for((i[0] * (i[1] / 8) + 3) / 4, 256, 148 * 8)), dim3(256), 0, st, pdl, a);
      break;
    }
    case TSR_E_COLSUM_FINALIZE:
      ce = launch_k(colsum_finalize_kernel, dim3((i[1] + 127) / 128), dim3(128), 0, st, pdl, (const float*)p[0], (float*)p[1], i[0], i[1], i[2], i[3],
                                                                i[4], i[5]);
      break;
    case TSR_E_SUM_FINALIZE:
      ce = launch_k(sum_finalize_kernel, dim3(1), dim3(32), 0, st, pdl, (const float*)p[0], (float*)p[1], i[0], i[1], d.f[0]);
      break;
    case TSR_E_PACK_W:
      ce = launch_k(pack_w_kernel, dim3(static_cast<unsigned>(i[1])), dim3(256), 0, st, pdl, (const tsr_pack_entry_t*)p[0], i[0]);
      break;
    case TSR_E_UNPACK_G:
      ce = launch_k(unpack_g_kernel, dim3(static_cast<unsigned>(i[1])), dim3(256), 0, st, pdl, (const tsr_pack_entry_t*)p[0], i[0]);
      break;
    case TSR_E_ADAM:
      ce = launch_k(adam_pack_kernel, dim3(static_cast<unsigned>(i[1])), dim3(256), 0, st, pdl,
                    (const tsr_adam_entry_t*)p[0], i[0], (const float*)p[1], d.f[0], d.f[1], d.f[2], d.f[3], d.f[4], d.f[5],
                    (float*)p[2], (int*)p[3]);
      break;
    case TSR_E_LINEAR_WGRAD: {
      dim3 grid((i[2] + 255) / 256, (i[1] + kLwF - 1) / kLwF);
      ce = launch_k(linear_wgrad_kernel, dim3(grid), dim3(256), kLwF * i[0] * sizeof(float), st, pdl, (const float*)p[0], (const float*)p[1],
                                                                         (float*)p[2], (float*)p[3], i[0], i[1], i[2]);
      break;
    }
    case TSR_E_LOSS:
      ce = launch_k(loss_kernel, dim3(static_cast<unsigned>(i[2])), dim3(256), 0, st, pdl, (const float*)p[0], (const float*)p[1], (float*)p[2],
                                                              (float*)p[3], i[0], i[1], d.f[0]);
      break;
    case TSR_E_PACK_GATHER:
      ce = launch_k(pack_gather_kernel, dim3(grid_for(i[0])), dim3(256), 0, st, pdl, (const float*)p[0], (const int*)p[1],
                    (bf16*)p[2], i[0]);
      break;
    case TSR_E_ZERO: {
      // a kernel (not a memset node): stays inside programmatic-launch chains and graph branches of kernel nodes
      if ((reinterpret_cast<uintptr_t>(p[0]) & 15) || (i[0] & 15)) return cudaMemsetAsync(p[0], 0, static_cast<size_t>(i[0]), st);
      ce = launch_k(zero_kernel, dim3(grid_for(i[0] / 16, 256, 148 * 8)), dim3(256), 0, st, pdl, (uint4*)p[0], i[0] / 16);
      break;
    }
    case TSR_E_UPSAMPLE2X:
      ce = launch_k(upsample2x_kernel, dim3(grid_for(i[0] * 4 * i[1] * i[2] * (i[3] / 8))), dim3(256), 0, st, pdl, (const bf16*)p[0], (bf16*)p[1],
                                                                                   i[0], i[1], i[2], i[3], i[4], i[5]);
      break;
    case TSR_E_UPSAMPLE2X_BWD:
      ce = launch_k(upsample2x_bwd_kernel, dim3(grid_for(i[0] * i[1] * i[2] * (i[3] / 8))), dim3(256), 0, st, pdl, (const bf16*)p[0], (bf16*)p[1],
                                                                                   i[0], i[1], i[2], i[3], i[4], i[5]);
      break;
    case TSR_E_HEAD:
      ce = launch_k(head_kernel, dim3(static_cast<unsigned>(i[0])), dim3(256), 0, st, pdl, (const float*)p[0], (const float*)p[1], (const float*)p[2],
                                                              (const float*)p[3], (float*)p[4], (float*)p[5], i[0], i[1],
                                                              i[2], d.f[0]);
      break;
    case TSR_E_HEAD_BWD:
      ce = launch_k(head_bwd_kernel, dim3((i[1] + 255) / 256), dim3(256), 0, st, pdl, (const float*)p[0], (const float*)p[1], (const float*)p[2],
                                                         (const float*)p[3], (float*)p[4], (bf16*)p[5], (float*)p[6],
                                                         (float*)p[7], (float*)p[8], i[0], i[1], i[2],
                                                         i[3] > 0 ? i[3] : i[1], d.f[0]);
      break;
    case TSR_E_AXPBY:
      ce = launch_k(axpby_kernel, dim3(grid_for(i[0] / 8)), dim3(256), 0, st, pdl, (const bf16*)p[0], (const bf16*)p[1], (bf16*)p[2], i[0], d.f[0],
                                                      d.f[1]);
      break;
    case TSR_E_MAXPOOL2:
      ce = launch_k(maxpool2_kernel, dim3(grid_for(i[0] * (i[1] / 2) * (i[2] / 2) * (i[3] / 8))), dim3(256), 0, st, pdl, (const bf16*)p[0],
                                                                                         (bf16*)p[1], i[0], i[1], i[2],
                                                                                         i[3]);
      break;
    case TSR_E_MAXPOOL2_BWD:
      ce = launch_k(maxpool2_bwd_kernel, dim3(grid_for(i[0] * (i[1] / 2) * (i[2] / 2) * (i[3] / 8))), dim3(256), 0, st, pdl, 
          (const bf16*)p[0], (const bf16*)p[1], (const bf16*)p[2], (bf16*)p[3], i[0], i[1], i[2], i[3], i[4]);
      break;
    case TSR_E_CAST:
      ce = launch_k(cast_kernel, dim3(grid_for(i[0])), dim3(256), 0, st, pdl, p[0], p[1], i[0], i[1]);
      break;
    case TSR_E_CROP_LR:
      if (i[1] > 128 || i[2] > 32 || i[1] % 4 || i[2] * 4 != i[1]) return cudaErrorInvalidValue;
      ce = launch_k(crop_lr_kernel, dim3(static_cast<unsigned>(i[0])), dim3(256), 0, st, pdl, (const uint8_t*)p[0],
                    (const long long*)p[1], (const int*)p[2], (const int*)p[3], (float*)p[4], (float*)p[5], i[1], i[2], i[3]);
      break;
    case TSR_E_FEAT_T:
      ce = launch_k(feat_t_kernel, dim3(grid_for(((i[0] + 63) / 64) * i[1] * i[2] * 64)), dim3(256), 0, st, pdl,
                    (const bf16*)p[0], (bf16*)p[1], i[0], i[1], i[2], i[3], i[4]);
      break;
    case TSR_E_GAN_LOSS:
      ce = launch_k(gan_loss_kernel, dim3(1), dim3(256), 0, st, pdl, (const float*)p[0], (const float*)p[1], (float*)p[2],
                    (float*)p[3], (float*)p[4], (const float*)p[5], i[0], i[1], i[2], d.f[0], d.f[1], d.f[2]);
      break;
    case TSR_E_AXPBY_F32:
      ce = launch_k(axpby_f32_kernel, dim3(grid_for((i[0] + 3) / 4)), dim3(256), 0, st, pdl, (const float*)p[0],
                    (const float*)p[1], (float*)p[2], (const float*)p[3], i[0], d.f[0], d.f[1]);
      break;
    case TSR_E_CHANSUM_NCHW: {
      dim3 grid(static_cast<unsigned>(i[1]), static_cast<unsigned>(i[3]));
      ce = launch_k(chansum_nchw_kernel, dim3(grid), dim3(256), 0, st, pdl, (const float*)p[0], (float*)p[1], i[0], i[1], i[2], i[3]);
      break;
    }
    default:
      return cudaErrorInvalidValue;
  }
  return ce != cudaSuccess ? ce : cudaGetLastError();
}

}  // namespace tsr
