// Parameter blocks shared by the host launch code and the tcgen05 implicit-GEMM kernels.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace tsr {

constexpr int kMaxTaps = 81;      // 9x9
constexpr int kBlockM = 128;      // UMMA M (TMEM lanes)
constexpr int kConvThreads = 320; // warp0 TMA, warp1 MMA/TMEM, warps 2..9 epilogue (2 per TMEM lane quadrant)
constexpr int kWgradThreads = 192; // warp0 TMA, warp1 MMA/TMEM, warps 2..5 epilogue
constexpr int kConvHeaderBytes = 15360;  // conv_igemm shared-memory header: barriers, epilogue scratch, column vectors
constexpr int kMaxBnGroups = 2;    // BatchNorm statistics groups per launch (real | fake halves of a discriminator batch)

enum OutMode : int {
  OUT_LINEAR = 0,     // off = n*os_n + ho*os_h + wo*os_w + col
  OUT_SHUFFLE = 1,    // PixelShuffle(2) store: column block (i,j) of 64 -> pixel (2ho+i, 2wo+j)
  OUT_UNSHUFFLE = 2,  // inverse: fine pixel (ho,wo) -> coarse pixel (ho/2,wo/2), column block ((ho&1)*2+(wo&1))
  OUT_GEMM_T_ATOMIC = 3, // split-K GEMM: atomicAdd(out[col*ld + row], v) in fp32
  OUT_GATHER_W = 4       // horizontal K-tap sum of the accumulator columns into fp32 NCHW (see EpiParams::gather_*)
};
enum Act : int { ACT_NONE = 0, ACT_PRELU = 1, ACT_LEAKY = 2, ACT_RELU = 3 };

struct EpiParams {
  void* out;             // bf16 or fp32
  void* out_preact;      // optional bf16 copy of the pre-activation value (same addressing as out)
  const float* bias;     // optional, indexed by global column
  const float* prelu;    // scalar slope (device) for ACT_PRELU / bwd_act == ACT_PRELU
  const void* res;       // optional bf16 residual, aux addressing:  v = (acc+bias)*acc_scale + res*res_scale + res2*res2_scale
  const void* res2;      // optional second residual (same addressing)
  const void* bwd_z;     // optional bf16 tensor, aux addressing: v *= act'(bwd_z)
  float* dalpha_partial; // optional [grid] partial sums of v*z*[z<0] (PReLU slope gradient)
  float* stats_partial;  // optional [stats_ld][2]: per-column sum / sum of squares of the stored value, accumulated
                         // with red.global.add (the caller zeroes it before the launch)
  int* err;              // watchdog / error flag
  long long* trace;      // optional per-CTA clock64 stamps [grid][8]
  long long os_n, os_h, os_w;     // out strides in elements
  long long aux_n, aux_h, aux_w;  // aux strides in elements
  int out_mode;
  int out_f32;           // 1: out is fp32
  int out_ch_off;        // added to column for out / preact
  int aux_ch_off;
  int n_valid;           // columns >= n_valid are not stored (multiple of 16)
  int act;               // applied last
  int bwd_act;           // ACT_PRELU: z holds pre-activation; ACT_LEAKY/ACT_RELU: z holds activation output
  int stats_ld;          // total columns in stats_partial rows
  int shuf_c;            // channels per sub-pixel block for OUT_SHUFFLE / OUT_UNSHUFFLE
  float acc_scale;
  float leaky_slope;
  float res_scale, res2_scale;
  int res_cols;          // residuals only touch columns < res_cols
  // Fused BatchNorm / activation backward reduction (data-gradient convs): the value v computed so far is the gradient
  // w.r.t. y = act(BN(x)); with z = x*scale+shift (scale/shift from bnr_coef, identity when null) the epilogue turns
  // it into dz = v * act'(z), stores dz, and accumulates per column sum(dz) and sum(dz*x) into stats_partial (and
  // sum(v*z*[z<=0]) into dalpha_partial) - the reductions bn_bwd_reduce_kernel would otherwise make in a second pass.
  // Split-K with in-kernel finalize (gridDim.z > 1, out_mode != OUT_GEMM_T_ATOMIC): every split adds its partial
  // tile into `ws` (fp32 [M][ws_ld], all zero between launches) with vector reductions and bumps the tile's counter;
  // the CTA that arrives last reads the sums back, re-zeroes them (and the counter) and runs the normal epilogue.
  float* ws;
  int* tile_counters;    // [tiles_m * tiles_n], zero between launches
  int ws_ld;
  const void* bnr_x;     // bf16 raw conv output x of the BatchNorm being differentiated (aux addressing), or null
  const float* bnr_coef; // [groups][4][bnr_c]: scale, shift, mean, invstd (forward coefficients) or null
  const float* bnr_prelu;
  int bnr_act;
  int bnr_c;
  // BatchNorm statistics groups: the reference calls the discriminator separately on the real and on the fake batch, so
  // each call normalises with its own batch statistics. Both batches run through one launch here; rows
  // [0, group_rows) are group 0, the rest group 1 (group_rows = 0: one group). group_rows is a multiple of 32, so a
  // warp's 32 accumulator rows always belong to one group. stats_partial is then [groups][stats_ld][2].
  int group_rows;
  // Fused training/eval BatchNorm forward (the conv feeds a BatchNorm): bnf_mode 1 = training - the epilogue accumulates
  // the column sums as usual, all CTAs of the launch meet at a grid barrier (bnf_counter, one arrival counter per N
  // tile, zero at launch; the host only selects this mode for grids that are co-resident), every CTA derives
  // scale/shift of its columns from the completed sums and applies  y = act(acc*scale + shift) + res*res_scale
  // straight from the TMEM accumulator; CTA 0 of each N tile publishes bnf_coef (scale, shift, mean, invstd per group)
  // and updates the running statistics (group 0 first, then group 1: two sequential momentum updates, as two calls
  // would). The raw accumulator is stored (bf16) to out_preact for backward. bnf_mode 2 = eval: scale/shift come from
  // the running statistics, no barrier, no statistics, nothing published.
  int bnf_mode;
  int bnf_c;                 // channels of the BatchNorm (row length of the coefficient arrays)
  unsigned int* bnf_counter; // [tiles_n]
  const float* bnf_gamma;
  const float* bnf_beta;
  float* bnf_rm;
  float* bnf_rv;
  long long* bnf_nbt;
  float* bnf_coef;           // [groups][4][bnf_c]
  long long bnf_count;       // rows behind each group's statistics
  float bnf_eps, bnf_momentum;
  // Fused BatchNorm backward APPLY on top of the fused reduction (bnr_x above; data-gradient convs of co-resident
  // grids): dz is parked in the TMEM accumulator instead of being stored, the CTAs meet at the grid barrier
  // (bnf_counter) once every column sum is complete, and each CTA turns its tile into the BatchNorm input gradient
  //   dx = A*dz + Bx*x + Cc,  A = gamma*invstd, Bx = -A*invstd*mean(dz*xhat), Cc = -A*(mean(dz) - mean*invstd*mean(dz*xhat))
  // (per statistics group) and stores only dx. The first CTA of each N tile publishes dgamma = sum(dz*xhat),
  // dbeta = sum(dz) (summed over the groups) and the PReLU slope gradient - what bn_bwd_apply_kernel did in a second
  // launch plus one HBM round trip of dz.
  int bnr_apply;
  void* bnr_dx;              // bf16 dx output, addressed like `out` (same strides / channel offset)
  const float* bnr_gamma;    // [bnr_c]
  float* bnr_dgamma;         // [bnr_c] or null
  float* bnr_dbeta;          // [bnr_c] or null
  float* bnr_dalpha;         // scalar out (= *dalpha_partial) or null
  long long bnr_count;       // rows per statistics group
  // OUT_GATHER_W (row-decomposed K x K conv with a handful of output channels, e.g. the generators' 9x9 64->3 output
  // conv): accumulator column kw*gather_c + c is the partial product of horizontal tap kw for channel c. The tile goes
  // through shared memory (fp32, odd row stride), each output position of the tile's span sums its gather_k shifted
  // columns and adds the result to out[n][c][h][w] (fp32 NCHW, zero at launch) with one atomic: positions whose taps
  // straddle two tiles receive two adds (commutative, so the result does not depend on the order); the tile that owns
  // the position adds the bias. Replaces a [M][block_n] fp32 round trip through HBM plus a gather kernel.
  const float* gather_bias;
  int gather_k, gather_pad, gather_c;
  int gather_rows;   // 2: every GEMM row (n, y2, x) holds the taps of the two output rows 2*y2 + r, r = column / 32 (ten
                     // vertical taps, traversal stride 2 along H, N = 64): twice the columns per fetched activation tile
  // nearest x2 of the result folded into the store: pixel (ho, wo) also goes to (2ho + i, 2wo + j) of out_rep2x
  void* out_rep2x;
  long long rep_n, rep_h, rep_w;
  int rep_ch_off;
};

struct ConvParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmA2;     // cluster_n == 2: im2col map with 64-pixel boxes (each CTA of the pair fetches half of the A box)
  CUtensorMap tmO[4];   // staged epilogue: output map(s), one per PixelShuffle sub-pixel block (linear stores use [0])
  EpiParams epi;
  int M_total;     // rows of the implicit GEMM (N*Ho*Wo traversal positions)
  int Ho, Wo;      // traversal grid (per image)
  int stride;      // traversal stride in input pixels (along H)
  int stride_w;    // ... along W (== stride unless the descriptor says otherwise)
  int lower_h, lower_w;
  int num_taps;
  int kc_per_tap;  // K chunks (of block_k) per tap
  int block_k;     // 16 / 32 / 64 elements
  int block_n;     // multiple of 16, <= 256
  int a_mode;      // 0: im2col NHWC; 1: 2D tiled, K-major rows; 2: 2D tiled, MN-major (A given as [K][M])
  int a_c0;        // first input channel
  int b_rows_per_tap;
  int b_chunk_rows;     // a_mode 1/2, > 0: K chunk j of the B operand starts at row j*b_chunk_rows, column 0 (chunked
                        // [chunks*rows][64] layout, see tsr_conv_desc_t.w_chunk_rows)
  int iters_per_split;  // K iterations handled per blockIdx.z
  int stages;
  int tmem_cols;
  int persistent;  // 0: one CTA per output tile; > 0: that many CTAs (per N tile) walk the M tiles with the weights
                   // of their N tile resident in shared memory (see conv_igemm.cu)
  uint32_t b_res_bytes;  // bytes of the resident weight region (0 when not persistent)
  uint32_t acc_cols;     // TMEM columns per accumulator stage
  // Halo mode (persistent kernel only; 3x3 stride-1 "same" convs): an M tile is halo_th output rows x halo_pw = W+2
  // positions of ONE image; per 64-channel chunk ONE tiled TMA load fetches the (halo_th+2) x halo_pw input patch
  // (zero-filled outside the image) and the nine taps are nine UMMA operands that start (dy*halo_pw + dx) rows into
  // that patch - instead of nine im2col loads of 128 separate 128-byte rows each. Outputs at the two halo positions of
  // every row are computed and discarded. tmA then holds the tiled map.
  // The image is cut into vertical strips of halo_sw = halo_pw - 2 output columns, so any image width works and the
  // patch stays small: a tile is halo_th rows of one strip, tiles of an image are numbered strip-fastest.
  int halo;              // 0 / 1
  int halo_th, halo_pw;  // output rows per tile, patch positions per row (strip width + 2)
  int halo_tiles_per_img;
  int halo_H, halo_W;    // image size (output == input size)
  int halo_strips;       // strips per image row: ceil(W / halo_sw)
  // generalised geometry (the 9 x 1 tap column of the row-decomposed 9x9 output conv uses the same machinery): the
  // patch is (halo_th + halo_eh) rows x halo_pw positions with origin (row0 + halo_lo_h, col0 + halo_lo_w), a tile
  // keeps halo_sw = halo_pw - halo_ew positions per row; 3x3: eh = ew = 2, lo = -1; 9x1: eh = 8, ew = 0, lo = (-4, 0)
  int halo_sw, halo_lo_h, halo_lo_w;
  // Staged epilogue (persistent FAST kernels, bf16 linear / PixelShuffle stores, 64-column N tiles): every epilogue warp
  // reads its whole share of the accumulator with one tcgen05.ld + wait, hands the TMEM stage back at once, and writes
  // the finished bf16 tile into a 128B-swizzled shared-memory buffer (double-buffered); one thread stores it with
  // cp.async.bulk.tensor (tmO) - full 128-byte lines instead of 32 half-sector stores per warp instruction. Halo tiles
  // are compacted (the two discarded positions per patch row are skipped), so the box is th x (pw-2) pixels.
  int staged;
  int res_reduce;        // staged epilogue, residual == output buffer (y = x + f(..) written in place, res_scale 1, no
                         // activation after the add): the tile is ADDED to global memory by the bulk store
                         // (cp.reduce.async.bulk.tensor .add) instead of loading x into registers first
  uint32_t extra_bytes;  // shared memory behind the operand ring: staging buffers (staged) or the OUT_GATHER_W tile
  // Activation multicast (one-tile FAST kernels, tiles_n even, M a multiple of 128): the launch runs clusters of two
  // CTAs along N - same M tile, neighbouring N tiles. Both need the same 128-pixel x 64-channel activation box per K
  // iteration: each fetches 64 pixels of it and multicasts them into both CTAs' ring stage, so the per-CTA L2 -> SM
  // traffic of an iteration drops from A + B to A/2 + B (the deep layers are bound by exactly that traffic). A stage is
  // recycled when both CTAs' MMA warps have consumed it (multicast tcgen05.commit onto both `empty` barriers).
  int cluster_n;   // 1 or 2
  int w_static;    // the B operand is not written by any kernel of the enclosing stream segment (real weights)
  int debug;       // attribution experiments only (TSR_CONV_DEBUG bits, tools/trace_conv.py), 0 in production: 1 = the
                   // epilogue skips the accumulator read-out and the stores, 2 = it computes but does not store,
                   // 4 = (persistent kernel) no activation TMA loads, 8 = (persistent kernel) no UMMAs issued.
  // derived on the host so that the single-thread producer / MMA loops stay short
  uint32_t a_bytes, b_bytes, stage_bytes;
  uint32_t ksteps;      // block_k / 16
  uint32_t sbo_bytes;   // 8 rows of the K-major tiles
  uint32_t layout_type; // UMMA swizzle code of the K-major tiles
  uint32_t idesc;
  uint16_t tap_off[kMaxTaps];   // (off_h << 8) | off_w
  uint16_t tap_wrow[kMaxTaps];  // tap slot in the packed weight matrix
};

constexpr int kMaxGroup = 4;
struct ConvGroup {
  ConvParams p[kMaxGroup];
  int tiles_n[kMaxGroup];
};

// Weight-gradient kernel:  dW[co][tap][ci] += sum_p X[pixel(p) + tap, ci] * dY[p, co]
// computed as D[(block, ci), co] with the "M blocks" enumerating (tap, ci_block) pairs.
struct WgradParams {
  CUtensorMap tmX;   // im2col NHWC map over the layer input (MN-major operand A, chan_block channels per box)
  CUtensorMap tmDy;  // 2D tiled map over dY [pixels, Cout] (MN-major operand B, dy_block channels per box)
  float* out;        // fp32 [Cout][num_taps][cin_pad], accumulated with red.global.add
  int* err;
  int M_total;       // pixels (traversal positions == rows of dY)
  int Ho, Wo;
  int stride;
  int lower_h, lower_w;
  int num_taps;
  int cin_blocks;      // cin_pad / chan_block
  int chan_block;      // channels per A box: 64 (SW128), 32 (SW64), 16 (SW32)
  int blocks_per_m;    // 128 / chan_block
  int groups_per_cta;  // UMMA M-groups (128 rows each) accumulated by one CTA
  int total_blocks;    // num_taps * cin_blocks
  int block_n;         // Cout tile
  int dy_block;        // channels per dY box: 64 / 32 / 16
  int cin_pad;
  int cout_valid;
  int pix_per_stage;   // 32 or 64
  int stages_per_cta;  // pipeline iterations per CTA (pixel range = stages_per_cta * pix_per_stage)
  int stages;          // smem ring depth
  int tmem_cols;
  uint16_t tap_off[kMaxTaps];
};

}  // namespace tsr
