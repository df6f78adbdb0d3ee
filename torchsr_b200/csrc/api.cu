// C-ABI of libtorchsr_b200.so (see include/torchsr_b200.h): descriptor validation, TMA tensor-map encoding,
// launch-parameter construction and the recorded-program executor. No torch dependency.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/torchsr_b200.h"
#include "conv_params.h"

namespace tsr {
cudaError_t launch_conv_igemm(const ConvParams& p, int tiles_n, int splits, cudaStream_t stream, bool pdl);
cudaError_t launch_conv_group(const ConvGroup& g, int n, cudaStream_t stream, bool pdl);
cudaError_t launch_conv_wgrad(const WgradParams& p, int gsets, int tiles_n, int splits, cudaStream_t stream, bool pdl);
cudaError_t launch_elt(const tsr_elt_desc_t& d, cudaStream_t st, bool pdl);
size_t conv_igemm_smem_bytes(const ConvParams& p);
int conv_igemm_max_coresident(const ConvParams& p, int* per_sm);
size_t conv_wgrad_smem_bytes(const WgradParams& p);
}  // namespace tsr

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;
int g_driver_version = 0;
int* g_watchdog = nullptr;  // device flag
std::mutex g_mu;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

int ensure_init() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_encode_tiled && g_encode_im2col && g_watchdog) return 0;
  cudaError_t ce = cudaFree(0);
  if (ce != cudaSuccess) return fail(-1, "CUDA runtime unavailable: %s", cudaGetErrorString(ce));
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (ce != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return fail(-2, "cuTensorMapEncodeTiled not found");
  g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  fn = nullptr;
  ce = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q);
  if (ce != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return fail(-2, "cuTensorMapEncodeIm2col not found");
  g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);
  cudaDriverGetVersion(&g_driver_version);
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return fail(-3, "torchsr_b200 requires an sm_100 GPU (found compute capability major %d)", major);
  ce = cudaMalloc(&g_watchdog, sizeof(int));
  if (ce != cudaSuccess) return fail(-1, "cudaMalloc failed: %s", cudaGetErrorString(ce));
  cudaMemset(g_watchdog, 0, sizeof(int));
  return 0;
}

CUtensorMapSwizzle swizzle_for_bytes(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 2D row-major bf16 matrix [rows][ld]; box = {box_cols, box_rows}
int encode_2d(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_cols * 2),
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(-10, "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d base=%p", (int)r,
                (long long)rows, (long long)cols, (long long)ld, box_cols, box_rows, base);
  return 0;
}

// NHWC bf16 tensor, im2col mode: box = {chan, pixels}
int encode_im2col(CUtensorMap* m, const void* base, int64_t N, int64_t H, int64_t W, int64_t C, int64_t ld, int lower_h,
                  int lower_w, int upper_h, int upper_w, int stride, int chan, int pixels, int stride_w = 0) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * W,
                           static_cast<cuuint64_t>(ld) * 2 * W * H};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride_w > 0 ? stride_w : stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                               upper, static_cast<cuuint32_t>(chan), static_cast<cuuint32_t>(pixels), estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(chan * 2),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(-11, "cuTensorMapEncodeIm2col failed (%d): N=%lld H=%lld W=%lld C=%lld ld=%lld lo=(%d,%d) up=(%d,%d) s=%d box=%dx%d",
                (int)r, (long long)N, (long long)H, (long long)W, (long long)C, (long long)ld, lower_h, lower_w, upper_h,
                upper_w, stride, chan, pixels);
  // Driver quirk for im2col maps over tensors smaller than 128 KiB on drivers <= 13.1 (descriptor bit 85 must be
  // clear); same fix-up CUTLASS applies when it builds im2col descriptors.
  if (g_driver_version <= 13010) {
    const uint64_t bytes = static_cast<uint64_t>(N) * H * W * ld * 2;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  }
  return 0;
}

int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

struct ConvLaunch {
  tsr::ConvParams p;
  int tiles_n, splits;
};
struct WgradLaunch {
  tsr::WgradParams p;
  int gsets, tiles_n, splits;
};

int g_use_persistent = -1;
bool use_persistent() {
  if (g_use_persistent < 0) {
    const char* e = getenv("TSR_CONV_PERSISTENT");
    g_use_persistent = (e && e[0] == '0') ? 0 : 1;
  }
  return g_use_persistent == 1;
}

// Halo mode (one patch per tile instead of nine im2col boxes: 1.4-1.75x instead of 9x re-read of the activations) only
// pays once the epilogue is cheaper than the main loop, i.e. together with the staged epilogue (conv_params.h): by
// default it is on exactly for the launches that get the staged epilogue. TSR_CONV_HALO=1 forces it for every eligible
// conv, TSR_CONV_HALO=0 switches it off; TSR_CONV_STAGED=0 switches the staged epilogue off. Read at descriptor-build
// time so tests can switch them.
int halo_setting() {   // -1 default, 0 off, 1 forced
  const char* e = getenv("TSR_CONV_HALO");
  if (!e || !e[0]) return -1;
  return e[0] == '1' ? 1 : 0;
}
bool use_staged() {
  const char* e = getenv("TSR_CONV_STAGED");
  return !(e && e[0] == '0') && use_persistent();
}

// persistent mode needs at least this many M tiles per CTA (x10); TSR_PERSIST_MIN_TILES_X10 overrides
int persistent_min_tiles_x10() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TSR_PERSIST_MIN_TILES_X10");
    v = e ? atoi(e) : 20;
    if (v < 10) v = 10;
  }
  return v;
}

// Activation multicast across clusters of two N tiles: parity-tested, but measured neutral (x0.94 - x1.04 on the VGG /
// discriminator layers with 128-512 input channels, tools/microbench_cluster.py: those launches already run at
// 0.9-1.2 PFLOP/s and are not bound by their L2 -> SM traffic), so it is opt-in: TSR_CONV_CLUSTER=1.
bool use_cluster() {
  const char* e = getenv("TSR_CONV_CLUSTER");
  return e && e[0] == '1';
}

int build_conv(const tsr_conv_desc_t& d, ConvLaunch* L, bool allow_persistent = true, bool allow_cluster = true) {
  using namespace tsr;
  if (int e = ensure_init()) return e;
  // the fused training BatchNorm keeps every tile's accumulator in TMEM across a grid barrier: one tile per CTA
  if (d.bnf_mode == 1 || d.bnr_apply) allow_persistent = false;
  ConvParams& p = L->p;
  memset(&p, 0, sizeof(p));
  if (d.block_k != 16 && d.block_k != 32 && d.block_k != 64) return fail(-20, "block_k must be 16/32/64");
  if (d.block_n < 16 || d.block_n > 256 || d.block_n % 16) return fail(-20, "block_n must be a multiple of 16 in [16,256]");
  if (d.cout_pad % d.block_n) return fail(-20, "cout_pad must be a multiple of block_n");
  if (d.num_taps < 1 || d.num_taps > kMaxTaps) return fail(-20, "num_taps out of range");
  if (d.a_mode == 0) {
    if (d.C % d.block_k) return fail(-20, "C (%lld) must be a multiple of block_k (%d)", (long long)d.C, d.block_k);
    const int stride_w = d.stride_w > 0 ? d.stride_w : d.stride;
    const int64_t wo = (d.W + d.upper_w - d.lower_w - 1) / stride_w + 1;
    const int64_t ho = (d.H + d.upper_h - d.lower_h - 1) / d.stride + 1;
    if (wo != d.Wo || ho != d.Ho)
      return fail(-21, "traversal grid mismatch: corners give %lldx%lld, descriptor says %lldx%lld", (long long)ho,
                  (long long)wo, (long long)d.Ho, (long long)d.Wo);
    if (int e = encode_im2col(&p.tmA, d.x, d.N, d.H, d.W, d.C, d.x_ld, d.lower_h, d.lower_w, d.upper_h, d.upper_w,
                              d.stride, d.block_k, kBlockM, stride_w))
      return e;
    p.M_total = static_cast<int>(d.N * d.Ho * d.Wo);
    p.kc_per_tap = static_cast<int>((d.C - d.a_c0) / d.block_k);
    p.Ho = static_cast<int>(d.Ho);
    p.Wo = static_cast<int>(d.Wo);
  } else if (d.a_mode == 1) {
    if (d.gemm_K % d.block_k) return fail(-20, "gemm_K must be a multiple of block_k");
    if (int e = encode_2d(&p.tmA, d.x, d.gemm_M, d.gemm_K, d.x_ld, d.block_k, kBlockM)) return e;
    p.M_total = static_cast<int>(d.gemm_M);
    p.kc_per_tap = static_cast<int>(d.gemm_K / d.block_k);
    p.Ho = p.Wo = 1;
  } else if (d.a_mode == 2) {
    if (d.block_k != 64) return fail(-20, "a_mode 2 requires block_k 64");
    if (d.gemm_K % d.block_k) return fail(-20, "gemm_K must be a multiple of block_k");
    if (int e = encode_2d(&p.tmA, d.x, d.gemm_K, d.gemm_M, d.x_ld, 64, d.block_k)) return e;
    p.M_total = static_cast<int>(d.gemm_M);
    p.kc_per_tap = static_cast<int>(d.gemm_K / d.block_k);
    p.Ho = p.Wo = 1;
  } else {
    return fail(-20, "bad a_mode");
  }
  if (d.w_chunk_rows > 0 && (d.a_mode == 0 || d.block_k != 64 || d.w_ld != 64 || d.w_chunk_rows < d.cout_pad))
    return fail(-20, "w_chunk_rows needs a_mode 1/2, block_k 64, w_ld 64 and w_chunk_rows >= cout_pad");
  if (int e = encode_2d(&p.tmB, d.w, d.w_rows, d.w_ld, d.w_ld, d.block_k, d.block_n)) return e;
  p.b_chunk_rows = d.w_chunk_rows;
  p.stride = d.stride;
  p.stride_w = d.stride_w > 0 ? d.stride_w : d.stride;
  p.lower_h = d.lower_h;
  p.lower_w = d.lower_w;
  p.num_taps = d.a_mode == 0 ? d.num_taps : 1;
  p.block_k = d.block_k;
  p.block_n = d.block_n;
  p.a_mode = d.a_mode;
  p.a_c0 = d.a_c0;
  p.w_static = d.w_static;
  {
    const char* dbg = getenv("TSR_CONV_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  // Staged epilogue (conv_params.h): plain bf16 linear / PixelShuffle stores of 64-column tiles from a persistent FAST
  // kernel. Decided here because its staging buffers come out of the shared-memory budget of the operand ring.
  const bool out_lin = d.out_mode == TSR_OUT_LINEAR, out_shuf = d.out_mode == TSR_OUT_SHUFFLE;
  const bool out_dense = d.os_h == d.Wo * d.os_w && d.os_n == d.Ho * d.os_h;
  const bool want_staged =
      use_staged() && allow_persistent && d.a_mode == 0 && (d.block_k == 64 || d.block_k == 32) && d.block_n == 64 && d.splits <= 1 &&
      (out_lin || (out_shuf && d.shuf_c == 64 && d.cout_pad == 256)) && !d.out_f32 && !d.out_preact && !d.bwd_z && !d.bnr_x &&
      !d.stats_partial && !d.dalpha_partial && !d.res2 && d.bnf_mode != 1 && !d.bnr_apply && !d.trace && p.debug == 0 &&
      !d.out_rep2x &&
      d.group_rows == 0 && d.os_w % 8 == 0 && d.os_h % 8 == 0 && d.os_n % 8 == 0 && d.out_ch_off % 8 == 0 &&
      reinterpret_cast<uintptr_t>(d.out) % 16 == 0 && d.n_valid % 8 == 0;
  const size_t smem_reserve = want_staged ? 2 * 16384 : 0;
  p.b_rows_per_tap = d.cout_pad;
  memcpy(p.tap_off, d.tap_off, sizeof(p.tap_off));
  memcpy(p.tap_wrow, d.tap_wrow, sizeof(p.tap_wrow));
  const int total_iters = p.num_taps * p.kc_per_tap;
  int splits = d.splits > 0 ? d.splits : 1;
  if (splits > 1 && d.out_mode != TSR_OUT_GEMM_T_ATOMIC && (!d.ws || !d.tile_counters || d.ws_ld < d.cout_pad))
    return fail(-20, "split-K needs OUT_GEMM_T_ATOMIC or a reduction workspace (ws, tile_counters, ws_ld >= cout_pad)");
  if (splits > total_iters) splits = total_iters;
  p.iters_per_split = (total_iters + splits - 1) / splits;
  splits = (total_iters + p.iters_per_split - 1) / p.iters_per_split;
  L->splits = splits;
  L->tiles_n = d.cout_pad / d.block_n;
  p.acc_cols = pow2_cols(d.block_n);
  p.tmem_cols = p.acc_cols;
  uint32_t stage_bytes = ((kBlockM * d.block_k * 2 + d.block_n * d.block_k * 2) + 1023u) & ~1023u;
  // ring budget of the one-tile kernels (KB; TSR_RING_KB overrides): 96 KB leaves room for a second CTA on the SM
  static const int ring_kb = [] {
    const char* e = getenv("TSR_RING_KB");
    const int v = e ? atoi(e) : 96;
    return v < 48 ? 48 : (v > 200 ? 200 : v);
  }();
  int stages = ring_kb > 96 ? 8 : 6;     // the barrier block holds eight stages
  while (stages > 2 && static_cast<size_t>(stages) * stage_bytes > static_cast<size_t>(ring_kb) * 1024) --stages;
  if (static_cast<size_t>(stages) * stage_bytes > 200 * 1024) return fail(-22, "tile does not fit shared memory");
  if (stages > p.iters_per_split) stages = p.iters_per_split < 1 ? 1 : p.iters_per_split;
  // Launches whose CTAs meet at a grid barrier (fused BatchNorm forward / backward apply) must be co-resident: grids
  // of more than one CTA per SM keep their ring small enough for two CTAs per SM (<= 96 KB each, conv_igemm.cu)
  if ((d.bnf_mode == 1 || d.bnr_apply) && d.a_mode == 0) {
    const long tiles = static_cast<long>((p.M_total + kBlockM - 1) / kBlockM) * L->tiles_n;
    if (tiles > 148)
      while (stages > 2 && 1024 + kConvHeaderBytes + static_cast<size_t>(stages) * stage_bytes > 96 * 1024) --stages;
  }
  // Persistent weight-stationary mode: many M tiles per CTA, the weights of the N tile resident in shared memory
  // (they are re-fetched by every CTA otherwise: half of the L2->SM traffic of the large-M layers), two accumulator
  // stages. Chosen when the resident weights fit and every CTA gets at least two tiles.
  p.persistent = 0;
  p.b_res_bytes = 0;
  {
    const int tiles_m = (p.M_total + kBlockM - 1) / kBlockM;
    const int n_ctas = L->tiles_n > 0 ? (148 / L->tiles_n > 0 ? 148 / L->tiles_n : 1) : 148;
    const uint32_t a_bytes = kBlockM * d.block_k * 2, b_bytes = static_cast<uint32_t>(d.block_n) * d.block_k * 2;
    const size_t b_res = static_cast<size_t>(total_iters) * b_bytes;
    const size_t budget = 227 * 1024 - 1024 - kConvHeaderBytes - 1024 - smem_reserve;
    if (use_persistent() && d.a_mode == 0 && splits == 1 && d.out_mode != TSR_OUT_GEMM_T_ATOMIC && allow_persistent &&
        b_res % 1024 == 0 && a_bytes % 1024 == 0 && b_res + 3 * static_cast<size_t>(a_bytes) <= budget &&
        10 * tiles_m >= persistent_min_tiles_x10() * n_ctas && 2 * p.acc_cols <= 512) {
      int st = static_cast<int>((budget - b_res) / a_bytes);
      if (st > 8) st = 8;
      p.persistent = n_ctas;
      p.b_res_bytes = static_cast<uint32_t>(b_res);
      p.tmem_cols = 2 * p.acc_cols;
      stage_bytes = a_bytes;
      stages = st;
    }
  }
  // Halo mode on top of the persistent kernel (see conv_params.h): stride-1 "same" convs with nine taps on 64-channel
  // chunks - 3x3 (offsets up to (2, 2)) and the 9x1 tap column of the row-decomposed 9x9 output conv (offsets (kh, 0)).
  p.halo = 0;
  const int halo_env = halo_setting();
  const bool gather_w = d.out_mode == TSR_OUT_GATHER_W;
  int eh = 0, ew = 0;    // extent of the tap offsets
  for (int t = 0; t < d.num_taps && t < kMaxTaps; ++t) {
    eh = std::max(eh, static_cast<int>(d.tap_off[t] >> 8));
    ew = std::max(ew, static_cast<int>(d.tap_off[t] & 0xFF));
  }
  if ((halo_env == 1 || (halo_env < 0 && (want_staged || gather_w))) && use_persistent() && allow_persistent && d.a_mode == 0 &&
      splits == 1 && d.out_mode != TSR_OUT_GEMM_T_ATOMIC && d.stride == 1 && p.stride_w == 1 && d.num_taps == 9 && d.Ho == d.H && d.Wo == d.W &&
      ((eh == 2 && ew == 2 && d.lower_h == -1 && d.lower_w == -1) || (eh == 8 && ew == 0 && d.lower_h == -4 && d.lower_w == 0)) &&
      d.block_k == 64 && (d.C - d.a_c0) % 64 == 0 && d.W >= 8 && d.H >= 4 && 2 * p.acc_cols <= 512) {
    bool taps_ok = true;
    // Strip geometry: a tile is th rows x (pw - ew) columns of one image, pw * th <= 128 accumulator rows. Pick the
    // patch width that wastes the fewest accumulator rows, weighing in the bytes of the patch it has to fetch. The
    // OUT_GATHER_W epilogue needs every warp's 32 accumulator rows to be 32 consecutive pixels: pw a multiple of 32.
    int pw = 0, th = 0, strips = 0, tiles_y = 0;
    {
      long best_tiles = 0;
      const int max_pw = d.W + ew < 128 ? static_cast<int>(d.W) + ew : 128;
      for (int cand = gather_w ? 32 : 6; cand <= (gather_w ? 128 : max_pw); cand += gather_w ? 32 : 1) {
        const int sw = cand - ew;
        int c_th = 128 / cand;
        if (c_th > d.H) c_th = static_cast<int>(d.H);
        if (c_th < 1 || sw < 1) continue;
        const int c_strips = static_cast<int>((d.W + sw - 1) / sw);
        const int c_ty = static_cast<int>((d.H + c_th - 1) / c_th);
        // cost ~ tiles x (accumulator rows + 0.15 x patch positions): tile count first, patch traffic second
        const long n_tiles = static_cast<long>(c_strips) * c_ty * (20 * 128 + 3 * (c_th + eh) * cand);
        if (pw == 0 || n_tiles <= best_tiles) {
          best_tiles = n_tiles;
          pw = cand;
          th = c_th;
          strips = c_strips;
          tiles_y = c_ty;
        }
      }
    }
    if (pw == 0) taps_ok = false;
    const uint32_t b_bytes = static_cast<uint32_t>(d.block_n) * 64 * 2;
    const size_t b_res = static_cast<size_t>(total_iters) * b_bytes;
    const uint32_t patch_tx = static_cast<uint32_t>((th + eh) * pw * 128);
    // the last tap's 128-row operand window starts eh * pw + ew rows into the stage
    const uint32_t patch_alloc = (static_cast<uint32_t>((128 + eh * pw + ew) * 128) + 1023u) & ~1023u;
    const size_t budget = 227 * 1024 - 1024 - kConvHeaderBytes - 1024 - smem_reserve;
    if (taps_ok && th >= 1 && b_res % 1024 == 0 && b_res + 2 * static_cast<size_t>(patch_alloc) <= budget) {
      const int tiles_per_img = strips * tiles_y;
      const int tiles = tiles_per_img * static_cast<int>(d.N);
      const int n_ctas = 148 / L->tiles_n > 0 ? 148 / L->tiles_n : 1;
      int st = static_cast<int>((budget - b_res) / patch_alloc);
      if (st > 4) st = 4;
      // the tiled map replaces the im2col map: box {64 channels, W+2 positions, th+2 rows, 1 image}
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.C), static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H),
                            static_cast<cuuint64_t>(d.N)};
      cuuint64_t strides[3] = {static_cast<cuuint64_t>(d.x_ld) * 2, static_cast<cuuint64_t>(d.x_ld) * 2 * d.W,
                               static_cast<cuuint64_t>(d.x_ld) * 2 * d.W * d.H};
      cuuint32_t box[4] = {64, static_cast<cuuint32_t>(pw), static_cast<cuuint32_t>(th + eh), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = g_encode_tiled(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d.x), dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(-12, "cuTensorMapEncodeTiled (halo patch) failed (%d)", (int)r);
      p.halo = 1;
      p.halo_th = th;
      p.halo_pw = pw;
      p.halo_sw = pw - ew;
      p.halo_lo_h = d.lower_h;
      p.halo_lo_w = d.lower_w;
      p.halo_tiles_per_img = tiles_per_img;
      p.halo_strips = strips;
      p.halo_H = static_cast<int>(d.H);
      p.halo_W = static_cast<int>(d.W);
      p.persistent = tiles < n_ctas ? tiles : n_ctas;
      p.b_res_bytes = static_cast<uint32_t>(b_res);
      p.tmem_cols = 2 * p.acc_cols;
      stage_bytes = patch_alloc;
      stages = st;
      p.stages = stages;
      p.a_bytes = patch_tx;
      p.b_bytes = b_bytes;
      p.stage_bytes = stage_bytes;
      p.ksteps = 4;
      p.sbo_bytes = 8 * 64 * 2;
      p.layout_type = 2u;
      p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(d.block_n >> 3) << 17) |
                (static_cast<uint32_t>(kBlockM >> 4) << 24);
      goto epilogue_params;
    }
  }
  p.stages = stages;
  p.a_bytes = kBlockM * d.block_k * 2;
  p.b_bytes = static_cast<uint32_t>(d.block_n) * d.block_k * 2;
  p.stage_bytes = stage_bytes;
  p.ksteps = d.block_k / 16;
  p.sbo_bytes = 8 * d.block_k * 2;
  p.layout_type = d.block_k == 64 ? 2u : (d.block_k == 32 ? 4u : 6u);
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((d.a_mode == 2 ? 1u : 0u) << 15) |
            (static_cast<uint32_t>(d.block_n >> 3) << 17) | (static_cast<uint32_t>(kBlockM >> 4) << 24);
epilogue_params:
  EpiParams& e = p.epi;
  e.out = d.out;
  e.out_preact = d.out_preact;
  e.bias = d.bias;
  e.prelu = d.prelu;
  e.res = d.res;
  e.bwd_z = d.bwd_z;
  e.dalpha_partial = d.dalpha_partial;
  e.stats_partial = d.stats_partial;
  e.err = g_watchdog;
  e.trace = reinterpret_cast<long long*>(d.trace);
  e.os_n = d.os_n;
  e.os_h = d.os_h;
  e.os_w = d.os_w;
  e.aux_n = d.aux_n;
  e.aux_h = d.aux_h;
  e.aux_w = d.aux_w;
  e.out_mode = d.out_mode;
  e.out_f32 = d.out_f32;
  e.out_ch_off = d.out_ch_off;
  e.aux_ch_off = d.aux_ch_off;
  e.n_valid = d.n_valid;
  e.act = d.act;
  e.bwd_act = d.bwd_act;
  e.stats_ld = d.stats_ld;
  e.shuf_c = d.shuf_c > 0 ? d.shuf_c : 64;
  e.acc_scale = d.acc_scale;
  e.leaky_slope = d.leaky_slope;
  e.res2 = d.res2;
  e.res_scale = d.res_scale;
  e.res2_scale = d.res2_scale;
  e.res_cols = d.res_cols > 0 ? d.res_cols : (1 << 30);
  e.ws = d.ws;
  e.tile_counters = d.tile_counters;
  e.ws_ld = d.ws_ld;
  e.bnr_x = d.bnr_x;
  e.bnr_coef = d.bnr_coef;
  e.bnr_prelu = d.bnr_prelu;
  e.bnr_act = d.bnr_act;
  e.bnr_c = d.bnr_c;
  e.group_rows = d.group_rows;
  e.bnf_mode = d.bnf_mode;
  e.bnf_c = d.bnf_c;
  e.bnf_counter = d.bnf_counter;
  e.bnf_gamma = d.bnf_gamma;
  e.bnf_beta = d.bnf_beta;
  e.bnf_rm = d.bnf_rm;
  e.bnf_rv = d.bnf_rv;
  e.bnf_nbt = reinterpret_cast<long long*>(d.bnf_nbt);
  e.bnf_coef = d.bnf_coef;
  e.bnf_count = d.bnf_count;
  e.bnf_eps = d.bnf_eps;
  e.bnf_momentum = d.bnf_momentum;
  if (e.group_rows < 0 || e.group_rows % 32) return fail(-20, "group_rows must be a non-negative multiple of 32");
  if (e.group_rows > 0 && (d.a_mode != 0 || e.group_rows >= p.M_total)) return fail(-20, "group_rows needs an im2col conv and 0 < group_rows < M");
  if (e.bnf_mode < 0 || e.bnf_mode > 2) return fail(-20, "bnf_mode must be 0, 1 or 2");
  if (e.bnf_mode) {
    if (d.a_mode != 0 || splits != 1 || d.out_mode != TSR_OUT_LINEAR || d.bias || d.bwd_z || d.bnr_x || d.res2 ||
        d.acc_scale != 1.f)
      return fail(-20, "fused BatchNorm forward needs an unsplit im2col conv with a linear store and no bias / bwd_z / bnr_x / res2 / acc_scale");
    if (e.bnf_c < e.n_valid || !e.bnf_rm != !e.bnf_rv) return fail(-20, "fused BatchNorm forward: bad bnf_c or running statistics");
    if (e.bnf_mode == 1 && (!e.stats_partial || !e.bnf_counter || e.bnf_count < 1 || p.persistent))
      return fail(-20, "training-mode fused BatchNorm needs stats_partial, bnf_counter, bnf_count and one tile per CTA");
    if (e.bnf_mode == 2 && (!e.bnf_rm || e.stats_partial || e.out_preact))
      return fail(-20, "eval-mode fused BatchNorm needs running statistics and takes no stats_partial / out_preact");
  }
  e.bnr_apply = d.bnr_apply;
  e.bnr_dx = d.bnr_dx;
  e.bnr_gamma = d.bnr_gamma;
  e.bnr_dgamma = d.bnr_dgamma;
  e.bnr_dbeta = d.bnr_dbeta;
  e.bnr_dalpha = d.bnr_dalpha;
  e.bnr_count = d.bnr_count;
  e.out_rep2x = d.out_rep2x;
  e.rep_n = d.rep_n;
  e.rep_h = d.rep_h;
  e.rep_w = d.rep_w;
  e.rep_ch_off = d.rep_ch_off;
  if (d.out_rep2x && (d.a_mode != 0 || d.out_mode != TSR_OUT_LINEAR || d.out_f32 || d.bnf_mode || d.bnr_apply || splits != 1 ||
                      d.rep_ch_off % 8 || d.rep_n % 8 || d.rep_h % 8 || d.rep_w % 8))
    return fail(-20, "out_rep2x needs an unsplit im2col conv with a linear bf16 store, no fused BatchNorm and 16-byte aligned strides");
  e.gather_bias = d.gather_bias;
  e.gather_k = d.gather_k;
  e.gather_pad = d.gather_pad;
  e.gather_c = d.gather_c;
  e.gather_rows = d.gather_rows == 2 ? 2 : 1;
  p.staged = 0;
  p.extra_bytes = 0;
  if (d.out_mode == TSR_OUT_GATHER_W) {
    if (d.a_mode != 0 || splits != 1 || L->tiles_n != 1 || !d.out_f32 || d.gather_k < 1 || d.gather_c < 1 ||
        d.gather_k * d.gather_c > d.block_n || d.gather_pad < 0 || d.gather_pad >= d.gather_k || d.out_preact || d.bias ||
        d.res || d.bwd_z || d.bnr_x || d.stats_partial || d.bnf_mode || (p.halo && p.halo_pw % 32))
      return fail(-20, "OUT_GATHER_W needs an unsplit im2col conv with one N tile, an fp32 NCHW output and a plain epilogue");
    const int g_rows = d.gather_rows == 2 ? 2 : 1;
    if (d.gather_k != 9 || d.gather_c != 3 || d.gather_pad != 4 || d.block_n != 32 * g_rows)
      return fail(-20, "OUT_GATHER_W is implemented for 9 taps x 3 channels in 32 columns per output row (conv_igemm.cu:gather9x3)");
    if (g_rows == 2 && (d.stride != 2 || p.stride_w != 1 || p.halo))
      return fail(-20, "OUT_GATHER_W with two output rows per GEMM row traverses H with stride 2 (im2col tiles)");
  }
  if (want_staged && p.persistent && (p.halo || (out_lin && out_dense))) {
    // output map(s): the staging buffer is one {64 channels x 128 pixels} (im2col tiles) or {64 x (pw-2) x th} (halo
    // tiles) box in the 128B swizzle; columns >= n_valid and pixels outside the tensor are clipped by the TMA unit
    const char* obase = reinterpret_cast<const char*>(d.out) + static_cast<int64_t>(d.out_ch_off) * 2;
    if (!p.halo) {
      if (int e2 = encode_2d(&p.tmO[0], obase, p.M_total, d.n_valid, d.os_w, 64, kBlockM)) return e2;
    } else {
      const int n_maps = out_shuf ? 4 : 1;
      for (int b = 0; b < n_maps; ++b) {
        const int64_t mul = out_shuf ? 2 : 1;
        const char* base = obase + ((b >> 1) * d.os_h + (b & 1) * d.os_w) * 2 * (out_shuf ? 1 : 0);
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(out_shuf ? 64 : d.n_valid), static_cast<cuuint64_t>(d.W),
                              static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.N)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(d.os_w * mul) * 2, static_cast<cuuint64_t>(d.os_h * mul) * 2,
                                 static_cast<cuuint64_t>(d.os_n) * 2};
        cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.halo_sw), static_cast<cuuint32_t>(p.halo_th), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = g_encode_tiled(&p.tmO[b], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims, strides,
                                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(-12, "cuTensorMapEncodeTiled (staged output) failed (%d)", (int)r);
      }
    }
    p.staged = 1;
    p.extra_bytes = 2 * 16384;
    // in-place residual (res aliases out element for element, added last with scale 1): the bulk store adds
    const bool res_inplace = d.res && out_lin && d.act == TSR_ACT_NONE && d.res_scale == 1.f &&
                             reinterpret_cast<const char*>(d.res) + static_cast<int64_t>(d.aux_ch_off) * 2 ==
                                 reinterpret_cast<const char*>(d.out) + static_cast<int64_t>(d.out_ch_off) * 2 &&
                             d.aux_n == d.os_n && d.aux_h == d.os_h && d.aux_w == d.os_w &&
                             (d.res_cols <= 0 || d.res_cols >= d.n_valid) && d.n_valid % 64 == 0;
    static const bool reduce_on = [] {
      const char* e = getenv("TSR_RES_REDUCE");
      return !(e && e[0] == '0');
    }();
    p.res_reduce = (res_inplace && reduce_on) ? 1 : 0;
  }
  // Activation multicast across clusters of two N tiles (conv_params.h): one-tile FAST kernels whose K loop is long
  // enough to be bound by the per-SM L2 -> SM traffic (at least 18 K iterations: 128 input channels), plain epilogues.
  p.cluster_n = 1;
  if (use_cluster() && allow_cluster && !p.persistent && d.a_mode == 0 && d.block_k == 64 && !d.trace && p.debug == 0 &&
      splits == 1 && L->tiles_n % 2 == 0 && p.M_total % kBlockM == 0 && d.bnf_mode != 1 && !d.bnr_apply &&
      total_iters >= 18) {
    if (int e2 = encode_im2col(&p.tmA2, d.x, d.N, d.H, d.W, d.C, d.x_ld, d.lower_h, d.lower_w, d.upper_h, d.upper_w,
                               d.stride, d.block_k, 64, p.stride_w))
      return e2;
    p.cluster_n = 2;
  }
  if (const char* v = getenv("TSR_CONV_VERBOSE"); v && v[0] == '1' && d.a_mode == 0)
    fprintf(stderr, "[tsr] conv M=%d C=%lld taps=%d block_n=%d tiles_n=%d persistent=%d halo=%d (th=%d pw=%d) staged=%d stages=%d "
            "stage_bytes=%u b_res=%u extra=%u out_mode=%d cluster=%d\n", p.M_total, (long long)d.C, p.num_taps, d.block_n, L->tiles_n,
            p.persistent, p.halo, p.halo_th, p.halo_pw, p.staged, p.stages, p.stage_bytes, p.b_res_bytes, p.extra_bytes, d.out_mode,
            p.cluster_n);
  if (e.bnr_apply && (!e.bnr_x || !e.bnr_dx || !e.bnr_coef || !e.bnr_gamma || !e.bnf_counter || e.bnr_count < 1 || splits != 1 ||
                      p.persistent || d.out_f32 || d.out_mode != TSR_OUT_LINEAR || e.bnf_mode))
    return fail(-20, "fused BatchNorm-backward apply needs bnr_x, bnr_coef, bnr_gamma, bnf_counter, bnr_count, a linear "
                     "bf16 store and one tile per CTA");
  if (e.bnr_x && !e.stats_partial) return fail(-20, "the fused BatchNorm-backward reduction needs stats_partial");
  if (e.bnr_x && e.bnr_act == TSR_ACT_PRELU && !e.bnr_prelu) return fail(-20, "bnr PReLU needs the slope pointer");
  if (e.bnr_x && e.bwd_z) return fail(-20, "bnr_x and bwd_z are mutually exclusive");
  if (!e.out) return fail(-20, "out is null");
  if ((e.act == TSR_ACT_PRELU || e.bwd_act == TSR_ACT_PRELU) && !e.prelu) return fail(-20, "PReLU needs the slope pointer");
  if (e.n_valid % 16) return fail(-20, "n_valid must be a multiple of 16");
  return 0;
}

int build_wgrad(const tsr_wgrad_desc_t& d, WgradLaunch* L) {
  using namespace tsr;
  if (int e = ensure_init()) return e;
  WgradParams& p = L->p;
  memset(&p, 0, sizeof(p));
  if (d.chan_block != 64 && d.chan_block != 32 && d.chan_block != 16) return fail(-30, "chan_block must be 16/32/64");
  if (d.dy_block != 64 && d.dy_block != 32 && d.dy_block != 16) return fail(-30, "dy_block must be 16/32/64");
  if (d.block_n % d.dy_block || d.block_n % 16 || d.block_n > 256) return fail(-30, "bad block_n");
  if (d.cout_valid % 16) return fail(-30, "cout_valid (accumulator row width) must be a multiple of 16");
  if ((d.C - d.x_c0) % d.chan_block) return fail(-30, "C must be a multiple of chan_block");
  if (d.num_taps < 1 || d.num_taps > kMaxTaps) return fail(-30, "num_taps out of range");
  const int64_t wo = (d.W + d.upper_w - d.lower_w - 1) / d.stride + 1;
  const int64_t ho = (d.H + d.upper_h - d.lower_h - 1) / d.stride + 1;
  if (wo != d.Wo || ho != d.Ho) return fail(-31, "traversal grid mismatch in wgrad descriptor");
  p.M_total = static_cast<int>(d.N * d.Ho * d.Wo);
  p.Ho = static_cast<int>(d.Ho);
  p.Wo = static_cast<int>(d.Wo);
  p.stride = d.stride;
  p.lower_h = d.lower_h;
  p.lower_w = d.lower_w;
  p.num_taps = d.num_taps;
  p.chan_block = d.chan_block;
  p.dy_block = d.dy_block;
  p.blocks_per_m = 128 / d.chan_block;
  p.cin_pad = static_cast<int>(d.C - d.x_c0);
  p.cin_blocks = p.cin_pad / d.chan_block;
  p.total_blocks = p.num_taps * p.cin_blocks;
  p.block_n = d.block_n;
  p.cout_valid = d.cout_valid;
  p.out = d.out;
  p.err = g_watchdog;
  memcpy(p.tap_off, d.tap_off, sizeof(p.tap_off));
  const int total_groups = (p.total_blocks + p.blocks_per_m - 1) / p.blocks_per_m;
  L->tiles_n = static_cast<int>((d.cout_valid + d.block_n - 1) / d.block_n);
  // Grid shape = (groups per CTA, pixels per stage, K splits), one CTA per SM (the accumulators may take the whole
  // TMEM). Every candidate is scored with a small cost model fitted to tools/microbench_wgrad.py on B200
  // (profiles/r02c_wgrad_grid.md) and the cheapest wins:
  //   pipeline stage: 300 clk + stage bytes / 36 B/clk  (the im2col boxes arrive at the per-SM L2 -> SM rate; the
  //                   UMMAs of the stage hide behind them), x1.5 when fewer than 3 stages fit in shared memory
  //   CTA:            3000 clk (setup, first TMA round trip, drain) + stages + tile bytes / 30 (fp32 vector atomics)
  //   launch:         max(1, CTAs / 148) rounds of that (short CTAs backfill, so fractional rounds), plus the total
  //                   atomic volume at ~12 KB/clk of L2 reduction throughput.
  // TSR_WGRAD_GPC / TSR_WGRAD_SPLITS / TSR_WGRAD_PIX pin a choice (tools/microbench_wgrad.py).
  auto stage_bytes = [&](int g, int px) {
    return static_cast<size_t>(((g * p.blocks_per_m * px * d.chan_block * 2 + (d.block_n / d.dy_block) * px * d.dy_block * 2) +
                                1023) & ~1023);
  };
  const size_t smem_cap = 200 * 1024;
  int gpc = 1, pix = 64, splits = 1;
  {
    const int env_gpc = getenv("TSR_WGRAD_GPC") ? atoi(getenv("TSR_WGRAD_GPC")) : 0;
    const int env_pix = getenv("TSR_WGRAD_PIX") ? atoi(getenv("TSR_WGRAD_PIX")) : 0;
    const int env_splits = getenv("TSR_WGRAD_SPLITS") ? atoi(getenv("TSR_WGRAD_SPLITS")) : 0;
    // measured: more than 3 groups per CTA never wins (two-stage rings, long atomics tail), and 32-pixel stages only
    // pay their doubled per-stage overhead when a 64-pixel stage does not fit at all
    const int max_gpc = std::min(env_gpc > 0 ? 8 : 3, std::min(512 / d.block_n, total_groups));
    double best = 1e30;
    for (int px = 64; px >= 32; px -= 32) {
      if (env_pix > 0 && px != env_pix) continue;
      if (env_pix <= 0 && px == 32 && best < 1e30) break;
      const int iters_total = (p.M_total + px - 1) / px;
      for (int g = 1; g <= max_gpc; ++g) {
        if (env_gpc > 0 && g != env_gpc) continue;
        if (stage_bytes(g, px) * 2 > smem_cap) continue;
        const int gsets = (total_groups + g - 1) / g;
        const int base = gsets * L->tiles_n;
        int st_fit = 6;
        while (st_fit > 2 && stage_bytes(g, px) * st_fit > smem_cap) --st_fit;
        const double it_clk = (300.0 + static_cast<double>(stage_bytes(g, px)) / 36.0) * (st_fit < 3 ? 1.5 : 1.0);
        const double tile_bytes = static_cast<double>(g) * 128 * d.block_n * 4;
        int last_ipc = -1;
        for (int sp = 1; sp <= iters_total; ++sp) {
          if (d.splits > 0 && sp != d.splits) continue;
          if (env_splits > 0 && sp != env_splits) continue;
          const int ipc = (iters_total + sp - 1) / sp;     // pipeline iterations per CTA
          if (ipc == last_ipc) continue;                   // same work per CTA with more CTAs: never better
          last_ipc = ipc;
          const int sp_eff = (iters_total + ipc - 1) / ipc;
          const long ctas = static_cast<long>(base) * sp_eff;
          const double rounds = std::max(1.0, static_cast<double>(ctas) / 148.0);
          const double cost = rounds * (3000.0 + ipc * it_clk + tile_bytes / 30.0) + ctas * tile_bytes / 12000.0;
          if (cost < best) {
            best = cost;
            gpc = g;
            pix = px;
            splits = sp_eff;
          }
        }
      }
    }
    if (best >= 1e30) return fail(-32, "wgrad tile does not fit shared memory");
  }
  p.groups_per_cta = gpc;
  p.pix_per_stage = pix;
  int stages = 6;
  while (stages > 2 && stage_bytes(gpc, pix) * stages > smem_cap) --stages;
  p.tmem_cols = pow2_cols(gpc * d.block_n);
  if (p.tmem_cols > 512) return fail(-32, "wgrad accumulators exceed TMEM");
  L->gsets = (total_groups + gpc - 1) / gpc;
  const int total_stage_iters = (p.M_total + pix - 1) / pix;
  if (splits > total_stage_iters) splits = total_stage_iters;
  p.stages_per_cta = (total_stage_iters + splits - 1) / splits;
  splits = (total_stage_iters + p.stages_per_cta - 1) / p.stages_per_cta;
  L->splits = splits;
  if (stages > p.stages_per_cta) stages = p.stages_per_cta;
  p.stages = stages;
  if (const char* v = getenv("TSR_CONV_VERBOSE"); v && v[0] == '1')
    fprintf(stderr, "[tsr] wgrad M=%d C=%d taps=%d N=%d block_n=%d: gpc=%d gsets=%d tiles_n=%d splits=%d pix=%d stages=%d/%d\n",
            p.M_total, p.cin_pad, p.num_taps, d.cout_valid, d.block_n, gpc, L->gsets, L->tiles_n, splits, pix, stages,
            p.stages_per_cta);
  if (int e = encode_im2col(&p.tmX, reinterpret_cast<const char*>(d.x) + static_cast<int64_t>(d.x_c0) * 2, d.N, d.H, d.W,
                            p.cin_pad, d.x_ld, d.lower_h, d.lower_w, d.upper_h, d.upper_w, d.stride, d.chan_block, pix))
    return e;
  if (int e = encode_2d(&p.tmDy, reinterpret_cast<const char*>(d.dy) + static_cast<int64_t>(d.dy_c0) * 2, p.M_total,
                        d.dy_c, d.dy_ld, d.dy_block, pix))
    return e;
  if (!p.out) return fail(-30, "out is null");
  return 0;
}

}  // namespace

struct tsr_prog {
  enum Kind { CONV, WGRAD, ELT, CONV_GROUP };
  struct Op {
    Kind kind;
    int side = 0;         // CONV: run on the weight-gradient side branch
    ConvLaunch conv;
    WgradLaunch wg;
    tsr_elt_desc_t elt;
    int group = -1;       // CONV_GROUP: index into groups
  };
  struct Group {
    tsr::ConvGroup g;
    int n;
  };
  std::vector<Group> groups;
  std::vector<Op> ops;
  // one instantiated CUDA graph per executed [first, last) range: every pointer in a program is fixed, so a range is
  // captured once and replayed with a single cudaGraphLaunch
  std::map<std::pair<int, int>, cudaGraphExec_t> graphs;
};

namespace {
cudaStream_t g_capture_stream = nullptr;
int g_use_graphs = -1;

bool use_graphs() {
  if (g_use_graphs < 0) {
    const char* e = getenv("TSR_GRAPHS");
    g_use_graphs = (e && e[0] == '0') ? 0 : 1;
  }
  return g_use_graphs == 1;
}

// Side branches for the weight-gradient GEMMs: per caller stream (two module backward passes captured on two
// streams must not serialise on one branch) a small pool of non-blocking streams used round-robin.
constexpr int kSideStreams = 2;
struct SideBranch {
  cudaStream_t s[kSideStreams] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[kSideStreams] = {nullptr, nullptr};
  // lowest-priority branch for bulk side work (elt.side == 2: the Linear weight gradient, thousands of small blocks):
  // its pending blocks only take the SM slots the main chain's kernels leave free
  cudaStream_t low = nullptr;
  cudaEvent_t join_low = nullptr;
};
std::map<cudaStream_t, SideBranch> g_side;
std::mutex g_side_mu;   // not g_mu: tsr_prog_run holds g_mu while it captures a range
int g_use_side = -1;

bool use_side_stream() {
  if (g_use_side < 0) {
    const char* e = getenv("TSR_WGRAD_BRANCH");
    g_use_side = (e && e[0] == '0') ? 0 : 1;
  }
  return g_use_side == 1;
}

SideBranch* side_for(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_side_mu);
  auto it = g_side.find(st);
  if (it != g_side.end()) return &it->second;
  SideBranch b;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = greatest priority
  const char* e = getenv("TSR_WGRAD_PRIO");
  const int prio = e ? atoi(e) : hi;
  for (int k = 0; k < kSideStreams; ++k) {
    if (cudaStreamCreateWithPriority(&b.s[k], cudaStreamNonBlocking, prio) != cudaSuccess ||
        cudaEventCreateWithFlags(&b.join[k], cudaEventDisableTiming) != cudaSuccess)
      return nullptr;
  }
  if (cudaEventCreateWithFlags(&b.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaStreamCreateWithPriority(&b.low, cudaStreamNonBlocking, lo) != cudaSuccess ||
      cudaEventCreateWithFlags(&b.join_low, cudaEventDisableTiming) != cudaSuccess)
    return nullptr;
  return &(g_side[st] = b);
}

int g_use_pdl = -1;
bool use_pdl() {
  if (g_use_pdl < 0) {
    const char* e = getenv("TSR_PDL");
    g_use_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_use_pdl == 1;
}

// Weight-gradient GEMMs only feed the final unpack, so they run on a side stream forked off the main chain right
// after the kernel that produced their dY (a parallel branch once the range is captured into a CUDA graph); the
// branch is joined once, before the op that consumes the accumulators (the unpack kernel / the end of the range).
//
// Programmatic dependent launch: a kernel on the main chain is launched with the PDL attribute when the operation
// right before it on that stream is another kernel of this range (not a memset, not the join of the side branch):
// its CTAs are then staged while the predecessor drains, and in a captured graph the edge becomes programmatic.
int launch_range(const tsr_prog* p, int first, int last, cudaStream_t st) {
  const bool side = use_side_stream();
  const bool pdl_on = use_pdl();
  SideBranch* sb = nullptr;
  if (side) {
    sb = side_for(st);
    if (!sb) return fail(-45, "side stream setup failed");
  }
  bool used[kSideStreams] = {false, false};
  bool used_low = false, prev_low = false;
  int next_side = 0;
  bool chain = false;   // previous op on `st` was one of our kernels
  auto join = [&]() -> cudaError_t {
    if (used_low) {
      used_low = false;
      chain = false;
      cudaError_t e = cudaEventRecord(sb->join_low, sb->low);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, sb->join_low, 0);
      if (e != cudaSuccess) return e;
    }
    for (int k = 0; k < kSideStreams; ++k) {
      if (!used[k]) continue;
      used[k] = false;
      chain = false;
      cudaError_t e = cudaEventRecord(sb->join[k], sb->s[k]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, sb->join[k], 0);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  bool prev_side_elt = false;   // consecutive side ELT ops stay on one branch, in order
  for (int i = first; i < last; ++i) {
    const tsr_prog::Op& op = p->ops[i];
    cudaError_t ce;
    const bool side_conv = side && op.kind == tsr_prog::CONV && op.side != 0;
    const bool side_elt = (side && op.kind == tsr_prog::ELT && op.elt.side != 0) || side_conv;
    if (side_elt && !side_conv && op.elt.side == 2) {
      ce = cudaSuccess;
      if (!prev_low) {
        ce = cudaEventRecord(sb->fork, st);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(sb->low, sb->fork, 0);
      }
      if (ce == cudaSuccess) ce = tsr::launch_elt(op.elt, sb->low, false);
      used_low = prev_low = true;
      prev_side_elt = false;
      if (ce != cudaSuccess) return fail(-42, "program op %d failed to launch: %s", i, cudaGetErrorString(ce));
      continue;
    }
    prev_low = false;
    if ((op.kind == tsr_prog::WGRAD && side) || side_elt) {
      int k = next_side;
      if (side_elt && prev_side_elt) {
        k = (next_side + kSideStreams - 1) % kSideStreams;   // same branch as the previous side op
        ce = cudaSuccess;
      } else {
        next_side = (next_side + 1) % kSideStreams;
        ce = cudaEventRecord(sb->fork, st);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(sb->s[k], sb->fork, 0);
      }
      if (ce == cudaSuccess)
        ce = side_conv ? tsr::launch_conv_igemm(op.conv.p, op.conv.tiles_n, op.conv.splits, sb->s[k], false)
             : side_elt ? tsr::launch_elt(op.elt, sb->s[k], false)
                        : tsr::launch_conv_wgrad(op.wg.p, op.wg.gsets, op.wg.tiles_n, op.wg.splits, sb->s[k], false);
      used[k] = true;
      prev_side_elt = side_elt;
      if (ce != cudaSuccess) return fail(-42, "program op %d failed to launch: %s", i, cudaGetErrorString(ce));
      continue;
    }
    prev_side_elt = false;
    if (op.kind == tsr_prog::WGRAD) {
      ce = tsr::launch_conv_wgrad(op.wg.p, op.wg.gsets, op.wg.tiles_n, op.wg.splits, st, pdl_on && chain);
      chain = true;
    } else {
      // the unpack kernel reads every accumulator: join the branches first
      if (op.kind == tsr_prog::ELT && op.elt.kind == TSR_E_UNPACK_G) {
        ce = join();
        if (ce != cudaSuccess) return fail(-45, "join failed: %s", cudaGetErrorString(ce));
      }
      if (op.kind == tsr_prog::CONV) {
        // experiment switch (profiles/r02c_two_ctas_per_sm.md): grid-barrier launches of more than one CTA per SM
        // without the programmatic-launch attribute
        static const bool big_barrier_pdl = [] {
          const char* e = getenv("TSR_PDL_BIG_BARRIER");
          return !(e && e[0] == '0');
        }();
        const tsr::ConvParams& cp = op.conv.p;
        const bool big_barrier = (cp.epi.bnf_mode == 1 || cp.epi.bnr_apply) && !cp.persistent &&
                                 static_cast<long>((cp.M_total + tsr::kBlockM - 1) / tsr::kBlockM) * op.conv.tiles_n > 148;
        ce = tsr::launch_conv_igemm(op.conv.p, op.conv.tiles_n, op.conv.splits, st,
                                    pdl_on && chain && (big_barrier_pdl || !big_barrier));
      } else if (op.kind == tsr_prog::CONV_GROUP) {
        ce = tsr::launch_conv_group(p->groups[op.group].g, p->groups[op.group].n, st, pdl_on && chain);
      } else {
        ce = tsr::launch_elt(op.elt, st, pdl_on && chain);
      }
      // an unaligned TSR_E_ZERO falls back to a memset node, which ends the programmatic chain
      chain = !(op.kind == tsr_prog::ELT && op.elt.kind == TSR_E_ZERO &&
                ((reinterpret_cast<uintptr_t>(op.elt.p[0]) & 15) || (op.elt.i[0] & 15)));
    }
    if (ce != cudaSuccess) return fail(-42, "program op %d failed to launch: %s", i, cudaGetErrorString(ce));
  }
  cudaError_t ce = join();
  if (ce != cudaSuccess) return fail(-45, "join failed: %s", cudaGetErrorString(ce));
  return 0;
}
}  // namespace

extern "C" {

int tsr_init(void) { return ensure_init(); }
const char* tsr_last_error(void) { return g_err.c_str(); }
int tsr_version(void) { return 1; }
int64_t tsr_launch_count(void) { return g_launches.load(); }

int tsr_conv(const tsr_conv_desc_t* d, void* stream) {
  ConvLaunch L;
  if (int e = build_conv(*d, &L)) return e;
  cudaError_t ce = tsr::launch_conv_igemm(L.p, L.tiles_n, L.splits, static_cast<cudaStream_t>(stream), false);
  if (ce != cudaSuccess) return fail(-40, "conv launch failed: %s", cudaGetErrorString(ce));
  g_launches++;
  return 0;
}

int tsr_conv_bnf_capacity(const tsr_conv_desc_t* d, int* ctas, int* capacity) {
  ConvLaunch L;
  if (int e = build_conv(*d, &L)) return e;
  const int tiles_m = (L.p.M_total + tsr::kBlockM - 1) / tsr::kBlockM;
  if (ctas) *ctas = L.p.persistent ? L.p.persistent * L.tiles_n : tiles_m * L.tiles_n * L.splits;
  int per_sm = 0;
  const int cap = tsr::conv_igemm_max_coresident(L.p, &per_sm);
  if (cap < 0) return fail(-46, "occupancy query failed");
  if (capacity) *capacity = cap;
  if (const char* v = getenv("TSR_CONV_VERBOSE"); v && v[0] == '1')
    fprintf(stderr, "[tsr] co-residency M=%d block_n=%d: %d CTAs, capacity %d (%d per SM, %zu B shared memory per CTA)\n",
            L.p.M_total, L.p.block_n, L.p.persistent ? L.p.persistent * L.tiles_n : tiles_m * L.tiles_n * L.splits, cap, per_sm,
            tsr::conv_igemm_smem_bytes(L.p));
  return 0;
}

int tsr_wgrad(const tsr_wgrad_desc_t* d, void* stream) {
  WgradLaunch L;
  if (int e = build_wgrad(*d, &L)) return e;
  cudaError_t ce = tsr::launch_conv_wgrad(L.p, L.gsets, L.tiles_n, L.splits, static_cast<cudaStream_t>(stream), false);
  if (ce != cudaSuccess) return fail(-40, "wgrad launch failed: %s", cudaGetErrorString(ce));
  g_launches++;
  return 0;
}

int tsr_elt(const tsr_elt_desc_t* d, void* stream) {
  if (int e = ensure_init()) return e;
  cudaError_t ce = tsr::launch_elt(*d, static_cast<cudaStream_t>(stream), false);
  if (ce != cudaSuccess) return fail(-41, "elementwise kind %d launch failed: %s", d->kind, cudaGetErrorString(ce));
  g_launches++;
  return 0;
}

tsr_prog_t* tsr_prog_create(void) { return new tsr_prog(); }
void tsr_prog_destroy(tsr_prog_t* p) {
  for (auto& kv : p->graphs) cudaGraphExecDestroy(kv.second);
  delete p;
}
int tsr_prog_size(const tsr_prog_t* p) { return static_cast<int>(p->ops.size()); }

int tsr_prog_add_conv(tsr_prog_t* p, const tsr_conv_desc_t* d) {
  tsr_prog::Op op;
  op.kind = tsr_prog::CONV;
  op.side = d->side;
  if (int e = build_conv(*d, &op.conv)) return e;
  p->ops.push_back(op);
  return static_cast<int>(p->ops.size()) - 1;
}
int tsr_prog_add_conv_group(tsr_prog_t* p, const tsr_conv_desc_t* descs, int n) {
  if (n < 1 || n > tsr::kMaxGroup) return fail(-20, "a conv group holds 1..%d members", tsr::kMaxGroup);
  tsr_prog::Group grp;
  memset(&grp.g, 0, sizeof(grp.g));
  grp.n = n;
  for (int k = 0; k < n; ++k) {
    ConvLaunch L;
    if (int e = build_conv(descs[k], &L, false, false)) return e;
    if (L.p.a_mode != 0 || L.splits != 1) return fail(-20, "conv group members must be unsplit im2col convs");
    if ((descs[k].bnr_apply != 0) != (descs[0].bnr_apply != 0) ||
        (descs[k].bnr_apply && k > 0 && (L.p.M_total != grp.g.p[0].M_total || L.tiles_n != grp.g.tiles_n[0])))
      return fail(-20, "a conv group with the fused BatchNorm-backward apply needs members of one tile grid (its grid "
                       "barrier counts every CTA of the launch)");
    grp.g.p[k] = L.p;
    grp.g.tiles_n[k] = L.tiles_n;
  }
  p->groups.push_back(grp);
  tsr_prog::Op op;
  op.kind = tsr_prog::CONV_GROUP;
  op.group = static_cast<int>(p->groups.size()) - 1;
  p->ops.push_back(op);
  return static_cast<int>(p->ops.size()) - 1;
}
int tsr_prog_add_wgrad(tsr_prog_t* p, const tsr_wgrad_desc_t* d) {
  tsr_prog::Op op;
  op.kind = tsr_prog::WGRAD;
  if (int e = build_wgrad(*d, &op.wg)) return e;
  p->ops.push_back(op);
  return static_cast<int>(p->ops.size()) - 1;
}
int tsr_prog_add_elt(tsr_prog_t* p, const tsr_elt_desc_t* d) {
  if (int e = ensure_init()) return e;
  tsr_prog::Op op;
  op.kind = tsr_prog::ELT;
  op.elt = *d;
  p->ops.push_back(op);
  return static_cast<int>(p->ops.size()) - 1;
}

int tsr_prog_run(tsr_prog_t* p, int first, int count, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = static_cast<int>(p->ops.size());
  const int last = count < 0 ? n : (first + count > n ? n : first + count);
  if (last <= first) return 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  if (!use_graphs() || last - first < 4 || cap != cudaStreamCaptureStatusNone) {
    // plain launches; when the caller's stream is itself being captured (a whole training step recorded into one
    // CUDA graph) the kernels, the memsets and the wgrad branch become nodes of the caller's graph
    if (int e = launch_range(p, first, last, st)) return e;
    g_launches += last - first;
    return 0;
  }
  const auto key = std::make_pair(first, last);
  auto it = p->graphs.find(key);
  if (it == p->graphs.end()) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_capture_stream) {
      cudaError_t ce = cudaStreamCreateWithFlags(&g_capture_stream, cudaStreamNonBlocking);
      if (ce != cudaSuccess) return fail(-43, "capture stream: %s", cudaGetErrorString(ce));
    }
    // first execution of this range: run it eagerly once (sets the kernels' function attributes outside capture and
    // keeps first-run behaviour identical), then capture it for every later call
    if (int e = launch_range(p, first, last, st)) return e;
    g_launches += last - first;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(g_capture_stream, cudaStreamCaptureModeThreadLocal);
    if (ce != cudaSuccess) return fail(-43, "begin capture: %s", cudaGetErrorString(ce));
    const int e = launch_range(p, first, last, g_capture_stream);
    ce = cudaStreamEndCapture(g_capture_stream, &graph);
    if (e) {
      if (graph) cudaGraphDestroy(graph);
      return e;
    }
    if (ce != cudaSuccess) return fail(-43, "end capture: %s", cudaGetErrorString(ce));
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail(-43, "graph instantiate: %s", cudaGetErrorString(ce));
    p->graphs[key] = exec;
    return 0;
  }
  cudaError_t ce = cudaGraphLaunch(it->second, st);
  if (ce != cudaSuccess) return fail(-44, "graph launch failed: %s", cudaGetErrorString(ce));
  g_launches += last - first;
  return 0;
}

int tsr_check_watchdog(void* stream) {
  if (int e = ensure_init()) return e;
  int v = 0;
  cudaError_t ce = cudaMemcpyAsync(&v, g_watchdog, sizeof(int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream));
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return fail(-50, "watchdog read failed: %s", cudaGetErrorString(ce));
  if (v != 0) {
    cudaMemsetAsync(g_watchdog, 0, sizeof(int), static_cast<cudaStream_t>(stream));
    return fail(v, "device watchdog fired: pipeline wait timed out (code %d)", v);
  }
  return 0;
}

}  // extern "C"
