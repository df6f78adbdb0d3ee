/*
 * torchsr_b200 C ABI — the drop-in boundary of the SRGAN/ESRGAN generator+discriminator hot path.
 *
 * The reference (roclark/torchsr) has no FFI: every FLOP is an ATen call made from nn.Module.forward and
 * autograd. This header is what the Python modules in torchsr_b200/ bind with ctypes in place
 * of those ATen calls; all pointers are raw device pointers, streams are cudaStream_t passed as void*, and
 * every entry point returns 0 or a negative error code (message via tsr_last_error()). No torch types.
 *
 * Entry point                      replaces (reference file:line)
 * -------------------------------  ---------------------------------------------------------------------------
 * tsr_conv / tsr_prog_add_conv     aten::convolution forward and its data gradient for every nn.Conv2d and
 *                                  nn.Linear on the path: srgan/generator.py:38,48,58  srgan/residual.py:27,64,67
 *                                  srgan/discriminator.py:31-69  esrgan/generator.py:34-52  esrgan/residual.py:33-52
 *                                  esrgan/discriminator.py:31-77; fused epilogues replace nn.PReLU/LeakyReLU
 *                                  (residual.py:29,66), nn.PixelShuffle (residual.py:28), the residual adds
 *                                  (residual.py:91, generator.py:78, esrgan/residual.py:86,129) and emit the
 *                                  BatchNorm batch statistics (residual.py:65,68).
 * tsr_wgrad / tsr_prog_add_wgrad   aten::convolution_backward (weight gradient) of the same layers.
 * tsr_elt / tsr_prog_add_elt       the HBM-bound remainder: layout conversion, BatchNorm finalize / apply /
 *                                  backward, activation backward, PixelShuffle/nearest-upsample index maps,
 *                                  loss reductions (srgan/trainer.py:163-165,384), weight pack / grad unpack,
 *                                  Linear weight gradient, classifier head (discriminator.py:64-69).
 * tsr_prog_*                       a recorded launch list with pre-encoded TMA descriptors (one per module
 *                                  call shape) replacing the per-op Python dispatch of nn.Sequential.forward.
 */
#ifndef TORCHSR_B200_H
#define TORCHSR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSR_MAX_TAPS 81

/* out_mode */
#define TSR_OUT_LINEAR 0
#define TSR_OUT_SHUFFLE 1
#define TSR_OUT_UNSHUFFLE 2
#define TSR_OUT_GEMM_T_ATOMIC 3
#define TSR_OUT_GATHER_W 4      /* horizontal tap sum into fp32 NCHW with atomics, see gather_* below */
/* activations */
#define TSR_ACT_NONE 0
#define TSR_ACT_PRELU 1
#define TSR_ACT_LEAKY 2
#define TSR_ACT_RELU 3

typedef struct tsr_conv_desc {
  /* A operand: a_mode 0 = NHWC bf16 activations read through TMA im2col; 1 = row-major [M][K] bf16 matrix;
     2 = row-major [K][M] bf16 matrix (M contiguous). */
  const void* x;
  const void* w;            /* packed bf16 weights, row-major [w_rows][w_ld]; rows = tap_slot*cout_pad + n */
  int64_t N, H, W, C;       /* a_mode 0: input tensor dims (C = channels visible to the map) */
  int64_t x_ld;             /* a_mode 0: pixel stride in elements; a_mode 1/2: row stride in elements */
  int64_t Ho, Wo;           /* a_mode 0: traversal grid per image */
  int64_t gemm_M, gemm_K;   /* a_mode 1/2 */
  int64_t w_rows, w_ld;
  int32_t a_mode;
  int32_t stride;           /* traversal stride */
  int32_t lower_h, lower_w; /* bounding-box lower corner (= -pad for a forward conv) */
  int32_t upper_h, upper_w; /* bounding-box upper corner (= pad - (k-1) for a forward conv) */
  int32_t num_taps;
  int32_t block_k;          /* 16, 32 or 64 channels per K chunk */
  int32_t block_n;          /* output columns per CTA: multiple of 16, <= 256 */
  int32_t cout_pad;         /* rows per tap slot in w (multiple of block_n) */
  int32_t a_c0;             /* first input channel used */
  int32_t splits;           /* split-K factor (OUT_GEMM_T_ATOMIC only) */
  uint16_t tap_off[TSR_MAX_TAPS];  /* (off_h << 8) | off_w, offsets relative to the lower corner */
  uint16_t tap_wrow[TSR_MAX_TAPS]; /* tap slot in w */
  /* epilogue */
  void* out;
  void* out_preact;
  const float* bias;
  const float* prelu;
  const void* res;
  const void* bwd_z;
  float* dalpha_partial;    /* one float accumulated atomically: PReLU slope gradient (pre-zeroed) */
  float* stats_partial;     /* [stats_ld][2] column sum / sum of squares, accumulated atomically (pre-zeroed) */
  int64_t os_n, os_h, os_w;
  int64_t aux_n, aux_h, aux_w;
  int32_t out_mode, out_f32, out_ch_off, aux_ch_off, n_valid, act, bwd_act, stats_ld, shuf_c;
  float acc_scale, leaky_slope;
  /* v = (acc + bias) * acc_scale + res * res_scale + res2 * res2_scale; res / res2 share the aux strides and only
     touch columns < res_cols (0 = all) */
  const void* res2;
  float res_scale, res2_scale;
  int32_t res_cols;
  int32_t w_static;         /* 1: `w` is not written by any kernel launched before this one on the stream since the
                               last full dependency (real, pre-packed weights): its tiles may be fetched before a
                               programmatically-launched kernel waits for its predecessor */
  /* fused BatchNorm / activation backward reduction (data-gradient convs, see csrc/conv_params.h): v is the gradient
     w.r.t. act(BN(x)); the epilogue stores dz = v * act'(x*scale+shift) and accumulates per column sum(dz), sum(dz*x)
     into stats_partial and the PReLU slope gradient into dalpha_partial. bnr_x uses the aux strides. */
  const void* bnr_x;
  const float* bnr_coef;    /* [4][bnr_c] forward coefficients (scale, shift, mean, invstd) or null (z = x) */
  const float* bnr_prelu;
  int32_t bnr_act, bnr_c;
  /* split-K with in-kernel finalize (splits > 1 and out_mode != TSR_OUT_GEMM_T_ATOMIC): fp32 workspace [M][ws_ld] and
     one int per output tile, both all-zero between launches (the kernel leaves them zeroed) */
  float* ws;
  int32_t* tile_counters;
  int32_t ws_ld;
  int32_t side;             /* inside a program: 1 = this GEMM only feeds gradient outputs, run it on the weight-gradient
                               side branch (as tsr_elt_desc_t.side) */
  int64_t* trace;           /* optional debug: per-CTA clock64 stamps [grid][40] (null in production) */
  /* BatchNorm statistics groups (one discriminator pass over the real | fake batches, srgan/trainer.py:446-447): rows
     [0, group_rows) are group 0, the rest group 1; 0 = one group. Multiple of 32. stats_partial / bnr_coef / bnf_coef
     then hold one block per group. */
  int32_t group_rows;
  /* fused BatchNorm forward (nn.BatchNorm2d after the conv, srgan/residual.py:65,68, discriminator.py:36-60):
     bnf_mode 1 = training (batch statistics, grid barrier inside the launch: needs stats_partial, bnf_counter and a
     co-resident grid - tsr_conv_bnf_capacity), 2 = eval (running statistics). out receives act(BN(acc)) + res,
     out_preact (optional) the raw bf16 conv output for backward. */
  int32_t bnf_mode;
  uint32_t* bnf_counter;    /* [cout_pad / block_n] arrival counters, zero at launch */
  const float* bnf_gamma;   /* [bnf_c] */
  const float* bnf_beta;
  float* bnf_rm;            /* running_mean, updated in training mode (may be null) */
  float* bnf_rv;
  int64_t* bnf_nbt;         /* num_batches_tracked, += number of groups */
  float* bnf_coef;          /* out, training mode: [groups][4][bnf_c] scale, shift, mean, invstd */
  int64_t bnf_count;        /* rows per statistics group */
  int32_t bnf_c;
  int32_t w_chunk_rows;     /* a_mode 1/2 only, > 0: w is stored as K chunks of 64 columns, chunk j = rows
                               [j*w_chunk_rows, (j+1)*w_chunk_rows) of a [chunks*w_chunk_rows][64] matrix (w_ld = 64): the
                               layout an all-gather of per-rank [rows][64] factor blocks produces */
  float bnf_eps, bnf_momentum;
  /* fused BatchNorm backward apply on top of bnr_x (nn.BatchNorm2d backward, autograd of srgan/residual.py:65,68,
     discriminator.py:36-60): the launch stores dx = A*dz + Bx*x + Cc instead of dz and publishes dgamma / dbeta /
     dalpha; needs a co-resident grid (tsr_conv_bnf_capacity), bnf_counter and bnr_coef. */
  int32_t bnr_apply, _pad3;
  void* bnr_dx;             /* bf16 BatchNorm input gradient, addressed with the out strides; `out` still receives dz when
                               bnr_act is NONE (dz is then the incoming gradient, which a skip path may still read) */
  const float* bnr_gamma;
  float* bnr_dgamma;
  float* bnr_dbeta;
  float* bnr_dalpha;
  int64_t bnr_count;
  /* TSR_OUT_GATHER_W (the 9x9 Cout=3 output conv of both generators, srgan/generator.py:58, row-decomposed): accumulator
     column kw*gather_c + c holds the partial product of horizontal tap kw for output channel c; the epilogue sums the
     gather_k horizontally shifted columns inside its tile and adds them to out (fp32 NCHW [N][gather_c][Ho][Wo], zero
     at launch):  out[n][c][h][w] += sum_kw acc[(n, h, w + kw - gather_pad)][kw*gather_c + c]  (+ gather_bias[c], once).
     One N tile (block_n == cout_pad >= gather_k*gather_c), no split-K; the out strides are not used. */
  const float* gather_bias;
  int32_t gather_k, gather_pad, gather_c;
  int32_t gather_rows;      /* 0 / 1: one output row per GEMM row. 2: the GEMM row (n, y2, x) produces the TWO output rows
                               2*y2 and 2*y2 + 1 (columns r*32 + kw*gather_c + c, ten vertical taps, traversal stride 2
                               along H): out is [N][gather_c][2*Ho][Wo], block_n = 64 */
  /* Nearest-neighbour x2 upsampling of the result (F.interpolate(scale_factor=2) in front of the ESRGAN upsample convs,
     esrgan/generator.py:73,76) folded into the producer's store: besides `out`, every output pixel (ho, wo) is written
     to the four positions (2ho + i, 2wo + j) of out_rep2x (bf16, strides in elements of the fine grid). Linear bf16
     stores of im2col convs only. */
  void* out_rep2x;
  int64_t rep_n, rep_h, rep_w;
  int32_t rep_ch_off;
  int32_t stride_w;         /* a_mode 0: traversal stride along W when it differs from `stride` (then the stride along H);
                               0 = same */
} tsr_conv_desc_t;

typedef struct tsr_wgrad_desc {
  const void* x;   /* layer input, NHWC bf16 */
  const void* dy;  /* output gradient, [pixels][dy_ld] bf16 */
  float* out;      /* fp32 [num_taps][cin_pad][cout_valid], pre-zeroed, accumulated with vector reductions */
  int64_t N, H, W, C, x_ld;
  int64_t Ho, Wo;
  int64_t dy_ld, dy_c;     /* dY row stride and channel count visible */
  int32_t stride, lower_h, lower_w, upper_h, upper_w;
  int32_t num_taps;
  int32_t chan_block;      /* 64 / 32 / 16 */
  int32_t dy_block;        /* 64 / 32 / 16 */
  int32_t block_n;
  int32_t cout_valid;      /* accumulator row width = padded output-channel count (multiple of 16) */
  int32_t x_c0, dy_c0;
  int32_t splits;          /* pixel splits (0 = auto) */
  uint16_t tap_off[TSR_MAX_TAPS];
} tsr_wgrad_desc_t;

/* Generic descriptor of the elementwise / reduction kernels; meaning of p/i/f per kind in csrc/eltwise.cu. */
typedef struct tsr_elt_desc {
  int32_t kind;
  int32_t side;    /* 1: inside a program this op only feeds gradient outputs - run it on the weight-gradient side
                      branch (consecutive side ops share one branch and keep their order) */
  void* p[12];
  int64_t i[16];
  float f[8];
} tsr_elt_desc_t;

enum tsr_elt_kind {
  TSR_E_IM2ROW = 1,
  TSR_E_GATHER_OUT = 2,
  TSR_E_NCHW2NHWC = 3,
  TSR_E_NHWC2NCHW = 4,
  TSR_E_BN_FINALIZE = 5,
  TSR_E_BN_EVAL_COEF = 6,
  TSR_E_BN_ACT = 7,
  TSR_E_BN_BWD_REDUCE = 8,
  TSR_E_BN_BWD_FINALIZE = 9,
  TSR_E_BN_BWD_APPLY = 10,
  TSR_E_ACT_BWD = 11,
  TSR_E_COLSUM_FINALIZE = 12,
  TSR_E_SUM_FINALIZE = 13,
  TSR_E_PACK_W = 14,
  TSR_E_UNPACK_G = 15,
  TSR_E_LINEAR_WGRAD = 16,
  TSR_E_LOSS = 17,
  TSR_E_ZERO = 18,
  TSR_E_UPSAMPLE2X = 19,
  TSR_E_UPSAMPLE2X_BWD = 20,
  TSR_E_HEAD = 21,
  TSR_E_HEAD_BWD = 22,
  TSR_E_AXPBY = 23,
  TSR_E_MAXPOOL2 = 24,
  TSR_E_MAXPOOL2_BWD = 25,
  TSR_E_CAST = 26,
  TSR_E_ADAM = 27,
  TSR_E_CHANSUM_NCHW = 28,
  TSR_E_GAN_LOSS = 29,   /* BCE / BCE-with-logits / relativistic-average GAN criteria, value + gradient, one launch */
  TSR_E_AXPBY_F32 = 30,  /* out = a * (*scalar) * x + b * y on fp32 vectors */
  TSR_E_CROP_LR = 32,    /* batched RandomCrop + flips + Pillow-exact bicubic /4 (dataset.py:86-99,118-121) on uint8 images in HBM */
  TSR_E_PACK_GATHER = 33, /* dst[i] (bf16) = idx[i] >= 0 ? src[idx[i]] (fp32) : 0; p: src, idx (int32), dst; i: 0 n. Operand
                             layouts that only one plan uses, re-derived from the fp32 parameter inside that plan's own
                             launch list (always current, also under CUDA-graph replays of optimizer steps) */
  TSR_E_FEAT_T = 31      /* NHWC bf16 features -> chunked transposed [(c,h,w)][batch] bf16 factor of the Linear wgrad GEMM */
};

/* weight pack / grad unpack index maps (TSR_E_PACK_W / TSR_E_UNPACK_G table entries) */
enum tsr_pack_mode {
  TSR_PK_FWD = 0,     /* dst[(kh*KW+kw)][co][ci]                       */
  TSR_PK_T = 1,       /* dst[(kh*KW+kw)][ci][co]   (data-gradient pack) */
  TSR_PK_ROWK = 2,    /* dst[kh][co][kw*Cin+ci]    (9x9, Cin=3 forward / its wgrad accumulator) */
  TSR_PK_ROWN = 3,    /* dst[kh][kw*Cout+co][ci]   (9x9, Cout=3 forward / its wgrad accumulator) */
  TSR_PK_ROWN_T = 4,  /* dst[kh][ci][kw*Cout+co]   (9x9, Cout=3 data gradient) */
  TSR_PK_FULLK = 5,   /* dst[0][co][(kh*KW+kw)*Cin+ci]  (3x3, Cin=3 forward / its wgrad accumulator) */
  TSR_PK_LINEAR = 6   /* dst[n][(h*Wf+w)*C+c] = src[n][c*Hf*Wf+h*Wf+w] */
};

typedef struct tsr_pack_entry {
  const void* src;   /* pack: fp32 OIHW parameter; unpack: fp32 accumulator in packed order */
  void* dst;         /* pack: bf16 packed; unpack: fp32 OIHW gradient */
  int32_t mode;
  int32_t cout, cin, kh, kw;
  int32_t rows_pad, cols_pad; /* packed matrix: rows per tap slot, columns */
  int32_t shuffle;   /* PixelShuffle(2) output-channel permutation: packed row r <-> co = 4*(r%(cout/4)) + r/(cout/4) */
  int64_t block_start;  /* first CUDA block of this entry (prefix sum, 256 threads x 4 elements per block) */
  int64_t count;        /* elements in the packed matrix */
} tsr_pack_entry_t;

/* TSR_E_ADAM table entry: one parameter tensor of a torch.optim.Adam-equivalent step (no weight decay, no amsgrad)
   that also refreshes the bf16 operand copies the conv / GEMM kernels read (csrc/eltwise.cu adam_pack_kernel). */
enum tsr_adam_mode {
  TSR_AD_PLAIN = 0,   /* update only (1-D parameters, convs whose packs the pack kernel still makes)       */
  TSR_AD_CONV = 1,    /* OIHW conv weight: + dst_fwd[(t*rows_fwd + co')*cols_fwd + ci], dst_t[(t*rows_t + ci)*cols_t + co'] */
  TSR_AD_LINEAR = 2,  /* Linear weight [n][c*HW+hw]: + dst_fwd[n*K + hw*C + c] (NHWC column order)          */
  TSR_AD_CONV_TILE = 3 /* as CONV for 3x3 weights with cin % 32 == 0, cout % 16 == 0, no PixelShuffle permutation:
                          one block = 16 x 32 x 9 tile, coalesced float4 state traffic, packs staged through smem */
};
typedef struct tsr_adam_entry {
  float* p;           /* parameter (fp32, updated in place) */
  const float* g;     /* gradient */
  float* m;           /* exp_avg */
  float* v;           /* exp_avg_sq */
  void* dst_fwd;      /* bf16 forward pack or null */
  void* dst_t;        /* bf16 data-gradient pack or null */
  int64_t numel;
  int64_t block_start; /* first CUDA block of this entry */
  int32_t mode;
  int32_t cout, cin, kk;        /* CONV: OIHW dims (kk = kh*kw); LINEAR: cout = rows, cin = C, kk = HW */
  int32_t rows_fwd, cols_fwd;   /* CONV: rows per tap slot / columns of the forward pack */
  int32_t rows_t, cols_t;       /* CONV: same for the transposed pack */
  int32_t shuffle;              /* PixelShuffle output-channel permutation (as tsr_pack_entry) */
  int32_t _pad;
} tsr_adam_entry_t;

typedef struct tsr_prog tsr_prog_t;

int tsr_init(void);
const char* tsr_last_error(void);
int tsr_version(void);

int tsr_conv(const tsr_conv_desc_t* d, void* stream);
/* number of CTAs of this conv's launch and how many the device can hold at once for its kernel / shared-memory
   footprint; the fused training BatchNorm (bnf_mode 1) requires ctas <= capacity. Returns 0 or a negative error. */
int tsr_conv_bnf_capacity(const tsr_conv_desc_t* d, int* ctas, int* capacity);
int tsr_wgrad(const tsr_wgrad_desc_t* d, void* stream);
int tsr_elt(const tsr_elt_desc_t* d, void* stream);

tsr_prog_t* tsr_prog_create(void);
void tsr_prog_destroy(tsr_prog_t* p);
int tsr_prog_add_conv(tsr_prog_t* p, const tsr_conv_desc_t* d);
/* n (<= 4) independent unsplit im2col convs executed by ONE launch (the output-parity classes of a stride-2 data
   gradient); counts as one op of the program */
int tsr_prog_add_conv_group(tsr_prog_t* p, const tsr_conv_desc_t* descs, int n);
int tsr_prog_add_wgrad(tsr_prog_t* p, const tsr_wgrad_desc_t* d);
int tsr_prog_add_elt(tsr_prog_t* p, const tsr_elt_desc_t* d);
int tsr_prog_size(const tsr_prog_t* p);
/* launches ops [first, first+count) on the stream; count < 0 means "to the end" */
int tsr_prog_run(tsr_prog_t* p, int first, int count, void* stream);
/* total kernels launched through this library since load (all entry points) */
int64_t tsr_launch_count(void);
/* reads and clears the device-side watchdog flag (0 = no kernel timed out); synchronises the stream */
int tsr_check_watchdog(void* stream);

#ifdef __cplusplus
}
#endif
#endif
