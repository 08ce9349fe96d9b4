/*
 * ffb200 -- C ABI of the B200-native FreqFusion x4 inference kernels (sm_100a).
 *
 * The reference (Nikhil-AI-Labs/image-super-resolution-2) is 100% PyTorch: there is no FFI in it.
 * Every entry point below replaces a group of torch operator calls on the hot path named in
 * BASELINE.json (`models/team29_FreqFusion/io.py:188 main` -> `CompleteEnhancedFusionSR.forward`);
 * the reference lines each one replaces are cited per function.  INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions:
 *   - plain pointers + explicit dims; the CALLER owns every buffer (device memory unless stated);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return 0 on success, negative on error (never throws/aborts); text via ff_last_error();
 *   - activations are NHWC ("pixel-major"): element (b,y,x,c) at ((b*H+y)*W+x)*ld + c;
 *   - bf16 buffers are `uint16_t`-sized; channel pitches (`*_ld`) are in ELEMENTS.
 */
#ifndef FFB200_H
#define FFB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFB200_ABI_VERSION 6

int ff_abi_version(void);
const char* ff_last_error(void);
/* Number of kernels launched by this library since load (bench.py reports it as gpu_launches). */
long long ff_launch_count(void);

/* activation codes used by the fused epilogues */
enum { FF_ACT_NONE = 0, FF_ACT_GELU = 1, FF_ACT_RELU = 2, FF_ACT_LRELU = 3, FF_ACT_SIGMOID = 4, FF_ACT_CLAMP01 = 5 };
/* convolution kinds of ff_conv_gemm */
enum { FF_CONV_1X1 = 0, FF_CONV_3X3 = 1, FF_CONV_2X2S2 = 2 };

/*
 * ff_conv_gemm -- implicit-GEMM convolution / linear layer on tcgen05 tensor cores.
 *   out[p, n] = epilogue( sum_{tap, c} x[pixel(p)+tap, c] * w[n, tap*cin + c] )
 * A tiles (128 output pixels = 8 rows x 16 cols) are fetched by 4-D TMA boxes shifted per tap with
 * hardware zero fill at the image border (= Conv2d zero padding); any H x W is accepted (edge tiles are partial: TMA clips the
 * stores, the direct-store epilogues mask by coordinates); B tiles by 2-D TMA; fp32
 * accumulators live in TMEM (double-buffered so the epilogue of tile i overlaps the MMAs of i+1).
 * Replaces: nn.Linear / nn.Conv2d call sites of hat_arch.py:156-158,83-85,65-69,596,874-893;
 * dat_arch.py:149-152,379-381,589-591,782; nafnet_arch.py:70-86,163-185; hierarchical_fusion.py:86-120;
 * enhanced_fusion.py:266-286; edge_enhancement.py:100-106,168-180 (incl. nn.PixelShuffle(2) folded
 * into the store, hat_arch.py:703, nafnet_arch.py:178-182).
 * Epilogue order:  v = acc + bias[n];  v = act(v);  v *= alpha;  v *= col_scale[n];  v *= mul[p,n];
 *                  v += aux_alpha * aux[p,n] * (aux_chan ? aux_chan[b,n] : 1);  v += res[p,n];
 *                  v = post_act(v);  store bf16 and/or fp32.
 */
typedef struct FFConvGemm {
  const void* x;       /* bf16 NHWC input */
  int B, H, W;         /* input batch / height / width */
  int x_ld;            /* channel pitch of x (multiple of 8) */
  int cin;             /* channels consumed per tap (multiple of 64, <= x_ld) */
  int kind;            /* FF_CONV_* */
  const void* w;       /* bf16 packed weights [n_pad][ntaps*cin] */
  int n_pad;           /* rows of w; multiple of 16 */
  int n_store;         /* columns written / read from the epilogue operands (<= n_pad) */
  const float* bias;   /* [n_pad] or NULL */
  int act;             /* FF_ACT_* applied right after bias */
  float alpha;         /* scalar scale after act */
  const float* col_scale; /* [n_pad] or NULL */
  const void* mul;     /* bf16 [pixels][mul_ld] or NULL */
  int mul_ld;
  const void* aux;     /* bf16 [pixels][aux_ld] or NULL */
  int aux_ld;
  const float* aux_chan; /* fp32 [B][aux_chan_ld] or NULL */
  int aux_chan_ld;
  float aux_alpha;
  const void* res;     /* residual [pixels][res_ld], fp32 if res_is_f32 else bf16, or NULL */
  int res_ld;
  int res_is_f32;
  int post_act;        /* FF_ACT_* applied after the residual add */
  void* out_bf16;      /* bf16 [pixels][out_ld] or NULL */
  int out_ld;
  float* out_f32;      /* fp32 [pixels][out_f32_ld] or NULL */
  int out_f32_ld;
  int pixel_shuffle;   /* 0, or 2: output is [B,2Ho,2Wo,n/4]; w rows must be packed as (i*2+j)*(n/4)+c */
  int gate_pairs;      /* 1: SimpleGate folded into the store: every 16-column chunk holds 8 x1 then 8 x2 channels,
                          out[p, n0/2 + i] = (acc+bias)[n0+i] * (acc+bias)[n0+8+i]; only bias + bf16 store apply */
  int w_batch_rows;    /* 0, or rows of w per sample: sample b uses weight rows [b*w_batch_rows, +n_pad) */
  int debug_simt;      /* 1: run the slow SIMT reference main loop (same epilogue) -- testing only */
  float* col_sums;     /* optional (plain bf16-store layers only): per-(128-pixel tile, 32-row quadrant) column sums of the stored
                        * values before bf16 rounding, [B][Ho*Wo/32][n_store] fp32 -- the global-average-pool partials of a
                        * squeeze-excite block, finished by ff_gap_finalize (deterministic two-phase sum) */
  /* Fused LayerNorm of the NEW residual-stream row (fp32-residual layers whose n tile spans the whole row, n_pad <= 256):
   * ln_out[p, n] = bf16((v[p, n] - mean_p) * rstd_p * ln_gamma[n] + ln_beta[n]) with the statistics over the first ln_cols
   * columns (columns ln_cols..n_store must be structurally zero: zero weights / bias / residual; their gamma / beta zero).
   * Replaces the separate nn.LayerNorm / LayerNorm2d pass that follows a residual add (hat_arch.py:268,308; dat_arch.py:727-733;
   * nafnet_arch.py:112,124): the row is still in TMEM when its statistics are known. */
  const float* ln_gamma; /* [n_store] or NULL */
  const float* ln_beta;  /* [n_store] */
  float ln_eps;
  int ln_cols;
  void* ln_out;          /* bf16 [pixels][ln_out_ld] or NULL */
  int ln_out_ld;
  /* Cropped output (0 = none): the layer computes H x W outputs but out_bf16 / out_f32 are [B][out_crop_h][out_crop_w] images
   * holding the top-left part -- the crop that follows an expert run on a padded image (expert_loader.py:612-615, 643-646,
   * nafnet_arch.py:216).  Direct-store epilogues only (n_store <= 4, or layers without per-pixel operands); `res` stays H x W. */
  int out_crop_h, out_crop_w;
  /* K-concatenated second operand of a 1x1 layer (NULL = none): out = [x | x2] . W^T with W [n_pad][cin + cin2]; x2 is a bf16 NHWC
   * tensor of the geometry of x.  With the weight block of x2 = diag(s_b) per sample (ff_build_concat_diag_weights,
   * w_batch_rows) an epilogue term  s_b[n] * x2[p, n]  runs on the tensor pipe instead (HAT: + 0.01 * cab * se, hat_arch.py:306). */
  const void* x2; int x2_ld; int cin2;
} FFConvGemm;
int ff_conv_gemm(const FFConvGemm* p, void* stream);
/* Per-sample weights [B][n_pad][k1 + n_pad] (bf16) for a K-concatenated layer: block 1 = w [n_pad][k1] (bf16, copied), block 2 =
 * diag(alpha * s[b][n]) -- see FFConvGemm.x2. */
int ff_build_concat_diag_weights(const void* w, int n_pad, int k1, const float* s, int s_ld, float alpha, int B, void* out, void* stream);

/*
 * ff_mlp_fused -- the transformer MLP of HAT's HAB / OCAB blocks as one kernel (csrc/mlp_fused.cu):
 *   x[p, :] += fc2(GELU(fc1(t[p, :])))          (Mlp.forward, hat_arch.py:77-94, called at :308 and :437)
 * t = the LayerNorm output (bf16, 192-wide rows: 180 channels + zero padding), fc1 192 -> 384 (360 + zero padding), fc2 384 -> 192.
 * The 128 x 384 hidden tile stays on the SM: fc1 accumulates 64 hidden columns at a time in TMEM, the GELU warps write them as a
 * bf16 A-operand tile in shared memory, fc2 accumulates over the six chunks.  x is the fp32 residual stream, updated in place;
 * out_bf16 (optional) receives a bf16 copy of the new x, ln_out (optional) LayerNorm(new x) for the next consumer (as in
 * FFConvGemm.ln_*).  Weights: packed bf16 w1 [384][192], w2 [192][384] (packing.pack_matrix), biases fp32 [384] / [192].
 */
typedef struct FFMlpFused {
  const void* t; int t_ld;       /* bf16 [B*H*W][t_ld] */
  int B, H, W;                   /* the token matrix as an image (tiles are 8 x 16 pixels) */
  const void* w1; const float* b1;
  const void* w2; const float* b2;
  float* x; int x_ld;            /* fp32 [B*H*W][x_ld], in place */
  void* out_bf16; int out_ld;    /* optional */
  const float* ln_gamma; const float* ln_beta; float ln_eps; int ln_cols;
  void* ln_out; int ln_out_ld;   /* optional */
} FFMlpFused;
int ff_mlp_fused(const FFMlpFused* p, void* stream);

/*
 * ff_png_encode_rgb8 -- host-side PNG writer of the plugin's save path (models/team29_FreqFusion/io.py:71-76 _save_image).
 * rgb: uint8 [h][w][3], row_stride bytes between rows; out: at least ff_png_bound_rgb8(h, w) bytes.  Returns the file size
 * (8-bit truecolour, Sub-filtered scanlines, one literal-only dynamic-Huffman deflate block) or a negative error code.
 * Pure CPU code, thread safe, no CUDA call: decoding the file gives back exactly the input pixels.
 */
long long ff_png_bound_rgb8(int h, int w);
long long ff_png_encode_rgb8(const unsigned char* rgb, int h, int w, long long row_stride, unsigned char* out, long long cap);

/*
 * ff_png_decode_rgb8 -- host-side PNG reader of the plugin's load path (models/team29_FreqFusion/io.py:64-68 _load_image:
 * Image.open(path).convert("RGB")).  file: the n bytes of a PNG file; rgb: uint8 [h][w][3] with cap >= h*w*3 bytes (null: only
 * out_h / out_w are filled from the header).  Handles 8-bit grey / grey+alpha / RGB / RGBA, non-interlaced (alpha dropped, grey
 * replicated -- what convert("RGB") does); returns 0, FF_PNG_UNSUPPORTED (1) for any other file (the caller falls back to its
 * general reader), negative on a bad argument.  Pure CPU code, thread safe; inflate is zlib's (libz.so.1 through dlopen).
 */
#define FF_PNG_UNSUPPORTED 1
int ff_png_decode_rgb8(const unsigned char* file, long long n, unsigned char* rgb, long long cap, int* out_h, int* out_w);

/*
 * ff_hab_tail -- everything of a HAT block after the attention as one kernel (hat_arch.py:303-309 HAB.forward tail,
 * :435-438 OCAB.forward tail):
 *     x1 = res + [a0 | a1] . wp^T + bp                  (attn.proj + shortcut; a1 / the second K half of wp carry the
 *                                                        0.01 * conv_block(x) * channel-attention term of :306 as a diagonal block)
 *     x  = x1 + fc2(GELU(fc1(LayerNorm(x1; ln2))))     (norm2 + Mlp + residual, :308)
 *     ln_out = LayerNorm(x; ln_gamma, ln_beta)          (optional: the norm1 of the next block)
 * x1 and LayerNorm(x1) never reach HBM (x1 stays in TMEM as the initial value of the fc2 accumulator).
 * a0, a1: bf16 [B*H*W][ld] with 192 columns (zero padded from 180); wp: bf16 [192 or B*192][K0] with K0 = 192 (a1 NULL) or 384,
 * per-sample rows when wp_batch_rows = 192 (see ff_build_concat_diag_weights); only the three 64 x 64 diagonal slabs of the
 * second K half are read.  res / x: fp32 [B*H*W][ld] (may alias).  w1 [384][192], w2 [192][384] bf16, biases fp32 padded.
 */
typedef struct FFHabTail {
  const void* a0; int a0_ld;
  const void* a1; int a1_ld;     /* optional */
  const float* a1_diag; int a1_diag_ld; float a1_alpha;   /* optional: the a1 term is a1[p][n] * bf16(a1_alpha * a1_diag[b][n]); the diagonal K block
                                   is generated on chip and wp is the shared [192][192] matrix (no per-sample weights in HBM) */
  int B, H, W;                   /* the token matrix as an image (tiles are 8 x 16 pixels) */
  const void* wp; int wp_batch_rows;
  const float* bp;
  const float* res; int res_ld;
  const float* ln2_gamma; const float* ln2_beta;   /* [192], zero padded */
  const void* w1; const float* b1;
  const void* w2; const float* b2;
  float* x; int x_ld;
  void* out_bf16; int out_ld;    /* optional bf16 copy of x */
  const float* ln_gamma; const float* ln_beta;
  void* ln_out; int ln_out_ld;   /* optional */
  float ln_eps; int ln_cols;     /* shared by both LayerNorms (180, 1e-5 in HAT) */
} FFHabTail;
int ff_hab_tail(const FFHabTail* p, void* stream);

/*
 * ff_naf_tail -- everything of a 64-channel NAFBlock that follows the SimpleGate depthwise conv, as ONE kernel (csrc/naf_tail.cu).
 * Replaces nafnet_arch.py:118-131 (x * sca -> conv3 -> y = inp + x * beta -> norm2 -> conv4 -> SimpleGate -> conv5 ->
 * y + x * gamma) and the norm1 of the following block (:112) at the 64-channel (full-resolution) levels, where the three
 * 1x1 convs are HBM-bound passes over the fp32 stream: one read of a0 and res, one write of x and out_bf16.
 *   y   = res + a0 . w3^T + b3                      w3 [64][64] bf16 with beta (and, per sample, sca) folded in by the caller:
 *                                                   w3[n][k] = beta[n] * conv3[n][k] * sca[b][k], rows b * w3_batch_rows + n; b3 = beta * bias
 *   t   = LayerNorm(y; ln2_gamma, ln2_beta)         over the 64 channels, biased variance, ln_eps
 *   u   = t . w4^T + b4                             w4 [128][64] bf16 in the reference's row order
 *   x   = y + (u[:, :64] * u[:, 64:]) . w5^T + b5   w5 [64][64] bf16 = gamma[n] * conv5[n][k], b5 = gamma * bias
 *   out_bf16 = LayerNorm(x; ln_gamma, ln_beta) when ln_gamma is given, else a bf16 copy of x (optional)
 * a0: bf16 [B*H*W][a0_ld]; res / x: fp32 [B*H*W][ld] (may alias); tiles are 8 x 16 pixels, any H x W.
 */
typedef struct FFNafTail {
  const void* a0; int a0_ld;
  int B, H, W;
  const void* w3; int w3_batch_rows;   /* 0 (one matrix) or 64 (one matrix per sample) */
  const float* b3;
  const float* res; int res_ld;
  const float* ln2_gamma; const float* ln2_beta;
  const void* w4; const float* b4;
  const void* w5; const float* b5;
  float* x; int x_ld;
  void* out_bf16; int out_ld;          /* optional */
  const float* ln_gamma; const float* ln_beta;   /* optional: out_bf16 receives the next LayerNorm instead of a copy */
  float ln_eps;
} FFNafTail;
int ff_naf_tail(const FFNafTail* p, void* stream);

/*
 * ff_window_attention -- fused window attention (QK^T + relative-position bias + shift mask + softmax + PV).
 * Replaces hat_arch.py:165-196 (WindowAttention.forward) with the roll / window_partition / window_reverse
 * copies of HAB.forward (hat_arch.py:279-303), OCAB's unfold + attention (hat_arch.py:398-435), and DAT's
 * SpatialAttention.forward (dat_arch.py:290-342) with the rolls of dat_arch.py:514-540.
 * Layout: qkv is bf16 [B*H*W][ld]; head h of q/k/v sits at channel {q,k,v}_off + (head_off+h)*32 (head dim 30
 * zero-padded to 32, q pre-scaled by head_dim^-0.5 * log2(e): the kernel's softmax uses exp2; padding dim 31 of v must hold 1.0 --
 * the softmax row sums are read from column 31 of P.V, so output channel 31 of every head comes out as 1.0 and dim 30 as 0).  Query window wh x ww must hold 256 tokens.  Key window kh x kw starts
 * kpad_{y,x} before the query window; keys outside the image are all-zero rows that still take softmax mass
 * (nn.Unfold zero padding, hat_arch.py:377).  Bias index for (query (qi,qj), key (ki,kj)) window coordinates:
 *   idx = (rel_sign*(qi-ki)+rel_off_y)*rel_stride + rel_sign*(qj-kj)+rel_off_x;  idx<0 -> idx+T  (HAT's OCA table
 *   relies on negative-index wrap-around, hat_arch.py:896-919; applied when rel_sign < 0);  bias = bias_table[(bias_head_off + h)*T + idx]
 *   (the table is stored transposed, [bias_heads][T]).  kw must be 8, 16, 24 or 32.
 * shift_{y,x} != 0 selects the cyclic shift and the {0,-100} region mask of hat_arch.py:921-940 / dat_arch.py:431-489.
 * Output token (un-shifted position) gets channels out_off + (head_off+h)*32 .. +32 of out (bf16 [B*H*W][out_ld]).
 */
typedef struct FFWinAttn {
  const void* qkv; int ld;
  int q_off, k_off, v_off;
  int B, H, W;
  int wh, ww, kh, kw, kpad_y, kpad_x;
  int shift_y, shift_x;
  int heads, head_off;
  const float* bias_table; int T, bias_heads, bias_head_off;
  int rel_sign, rel_off_y, rel_off_x, rel_stride;
  void* out; int out_ld, out_off;
  int Hp, Wp;   /* 0, or the padded extent (multiples of wh / ww, >= H / W): windows, shift and mask regions are laid on the
                 * Hp x Wp grid and tokens beyond H x W are all-zero q / k / v rows (DAT's F.pad of the projected qkv,
                 * dat_arch.py:505-512); their outputs are dropped (the crop of :556-557) */
} FFWinAttn;
int ff_window_attention(const FFWinAttn* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Memory-bound building blocks (csrc/pointwise.cu).  All NHWC.
 * ------------------------------------------------------------------------------------------------ */

/* LayerNorm over the channel axis of rows [rows][in_ld] (first C channels), one warp per row, fp32 math.
 * Writes bf16 and/or fp32; columns C..out_cols-1 are written as zero (channel padding of the GEMM operands).
 * Replaces nn.LayerNorm (hat_arch.py:232,256,378,391; dat_arch.py:109,686,711) and LayerNorm2d
 * (nafnet_arch.py:26-41, eps 1e-6, biased variance) -- identical in NHWC. */
int ff_layernorm(const void* x, int x_is_bf16, int in_ld, long long rows, int C, const float* gamma, const float* beta,
                 float eps, void* out_bf16, int out_ld, int out_cols, float* out_f32, int out_f32_ld, void* stream);

/* Global average pool over the P pixels of each sample: x [B][P][ld] -> out [B][out_ld] (fp32), two-phase
 * deterministic reduction through `scratch` (>= B*64*C floats).  nn.AdaptiveAvgPool2d(1) of hat_arch.py:50,
 * dat_arch.py:411,603, nafnet_arch.py:86. */
int ff_gap(const void* x, int x_is_bf16, int ld, int B, int P, int C, float* out, int out_ld, float* scratch,
           size_t scratch_bytes, void* stream);
/* Second phase of the pool on its own: out[b][c] = inv * sum_s partial[b][s][c] (partials from FFConvGemm.col_sums). */
int ff_gap_finalize(const float* partial, int B, int nsplit, int C, float inv, float* out, int out_ld, void* stream);

/* y[r][n] = act(x[r][:K] . W[n][:K] + bias[n]) for per-sample vectors (SE / SCA / channel-interaction heads);
 * columns N..y_cols-1 are written as zero. */
int ff_vec_linear(const float* x, int x_ld, int R, int K, const float* W, const float* bias, int N, int act, float* y,
                  int y_ld, int y_cols, void* stream);

/* ff_gap_finalize fused with the per-sample MLP that consumes the pooled vector (one launch instead of three):
 *   mean[b][c]   = inv * sum_s partial[b][s][c]                       (also written out: [B][mean_ld], columns C.. zero)
 *   h1 > 0:  hid = act1(w1 [h1][k1] . mean[:k1] + b1),  out[b][n] = act2(w2 [n_out][h1] . hid + b2)
 *   h1 == 0: out[b][n] = act1(w1 [n_out][k1] . mean[:k1] + b1)
 * columns n_out..out_cols-1 of out are written as zero.  counters: B zero-initialised uint32 (left zero by every launch).
 * Squeeze-excite of HAT's CAB (hat_arch.py:45-58), DAT's channel interaction (dat_arch.py:411-416), NAFNet's SCA (nafnet_arch.py:86-89). */
int ff_gap_finalize_mlp(const float* partial, int B, int nsplit, int C, float inv, float* mean, int mean_ld, unsigned int* counters,
                        const float* w1, const float* b1, int k1, int h1, int act1, const float* w2, const float* b2, int n_out, int act2,
                        float* out, int out_ld, int out_cols, void* stream);

/* Depthwise kh x kw convolution (zero padding), bf16 NHWC in/out, fp32 weights [kh*kw][C] and accumulate.
 * mode 0: out = act(dw(x)+bias) * (mul ? mul : 1);  mode 1 (SimpleGate, nafnet_arch.py:47-52,112-114):
 * out[c] = (dw(x)+bias)[c] * (dw(x)+bias)[c + C/2].  Also dat_arch.py:115,403-407 and the LKA chain
 * (large_kernel_attention.py:59-78). */
int ff_dwconv(const void* x, int x_ld, int B, int H, int W, int C, int kh, int kw, const float* w, const float* bias,
              int act, int mode, const void* mul, int mul_ld, void* out, int out_ld, void* stream);

/* 3x3 depthwise conv (same operands as ff_dwconv) that also emits the global-average-pool partials of its output:
 * col_sums [B][rows][cout] fp32 with rows = ff_dwconv_pool_rows(H, W, cout, mode) tile sums of the stored values before
 * bf16 rounding; ff_gap_finalize(col_sums, B, rows, cout, 1/(H*W), ...) finishes the pool.  Fuses the AdaptiveAvgPool2d(1)
 * of nafnet_arch.py:86,117 (SCA on the SimpleGate output) and dat_arch.py:411 (channel interaction on the conv branch)
 * into the producer.  ff_dwconv_pool_rows returns 0 when the shape does not tile (use ff_dwconv + ff_gap then). */
int ff_dwconv_pool(const void* x, int x_ld, int B, int H, int W, int C, const float* w, const float* bias, int act, int mode,
                   const void* mul, int mul_ld, void* out, int out_ld, float* col_sums, void* stream);
int ff_dwconv_pool_rows(int H, int W, int cout, int mode);

/* x[p][c] *= s[b][c] in place (bf16): NAFNet simplified channel attention, nafnet_arch.py:118. */
int ff_scale_channels(void* x, int ld, int B, long long pixels_per_sample, int C, const float* s, int s_ld, void* stream);

/* Per-sample weights with a channel-attention scale folded into the input columns: out[b][n][k] = bf16(w[n][k] * s[b][k]),
 * zero outside [N) x [K), laid out [B][n_pad][k_pad] for FFConvGemm.w_batch_rows = n_pad.  Replaces the in-place
 * `x * sca(x)` of nafnet_arch.py:118 when the weight is smaller than the activation (conv3(x * s_b) = conv3_b(x)). */
int ff_scale_weight_cols(const float* w, int N, int K, const float* s, int s_ld, int B, void* out, int n_pad, int k_pad,
                         void* stream);

/* fp32 -> split-bf16 operand packing, so fp32 layers of the routing path / image-space first layers run on ff_conv_gemm:
 * out[p][(t*k*k + tap)*Cin + c] = term_t(x[p + tap][c]), terms (hi, lo, hi) with hi = bf16(x), lo = bf16(x - hi); zero outside
 * the image and in the padding up to the next multiple of 64 columns.  Pair with weight rows [w_hi; w_hi; w_lo] (terms = 3,
 * ~16-bit mantissas on both operands, fp32 accumulation) or [w_hi; w_hi] (terms = 2).  k = 3 gathers the 3x3 neighbourhood
 * (im2col) so a 3-channel 3x3 conv becomes one 64-wide k-block.  Instantiated: (Cin,k,terms) = (3,3,2), (64,1,3), (32,1,3).
 * Replaces the fp32 nn.Conv2d evaluation of fusion_network.py:167-236 (difficulty / gate heads) and the 3->64 first layers. */
int ff_pack_taps(const float* x, int x_ld, int B, int H, int W, int Cin, int k, int terms, void* out, int out_ld, void* stream);

/* Direct fp32 convolution (k = 1 or 3, zero pad) for small channel counts: image-space first/last layers and
 * the fp32 routing path of the fusion head (fusion_network.py:167-236,543-607).  w is fp32 [Cout_pad][k*k*Cin]. */
int ff_conv_direct(const void* x, int x_is_bf16, int x_ld, int B, int H, int W, int Cin, int k, const float* w,
                   const float* bias, int Cout_pad, int n_store, int act, const float* mul_f32, int mul_ld,
                   void* out_bf16, int out_ld, float* out_f32, int out_f32_ld, void* stream);

/* Same as ff_nchw_to_nhwc into a right / bottom padded image [B*Hp*Wp][ld]: mode 1 = reflect padding (pad_to_window_size,
 * expert_loader.py:63-91: F.pad(..., mode='reflect') up to the next multiple of the window size), mode 0 = zero padding. */
int ff_nchw_to_nhwc_pad(const float* x, int B, int C, int H, int W, const float* sub, float* out, int ld, int Hp, int Wp, int mode, void* stream);

/* Layout conversion at the boundary: NCHW fp32 image <-> NHWC fp32 rows (optional per-channel subtraction,
 * hat_arch.py:972-973 `(x - mean) * img_range`). */
int ff_nchw_to_nhwc(const float* x, int B, int C, int H, int W, const float* sub, float* out, int ld, void* stream);
int ff_nhwc_to_nchw(const float* x, int ld, int coff, int B, int C, int H, int W, float* out, void* stream);

/* Bicubic up-sampling (a = -0.75, align_corners=False, border-clamped taps), NCHW fp32 -> NHWC fp32:
 * F.interpolate(mode='bicubic') of nafnet/__init__.py:128-133. */
int ff_bicubic_up(const float* x, int B, int C, int h, int w, int scale, float* out, int ld, void* stream);
/* ... into an image zero-padded on the right / bottom to Hp x Wp (NAFNet.check_image_size, nafnet_arch.py:219-225). */
int ff_bicubic_up_pad(const float* x, int B, int C, int h, int w, int scale, float* out, int ld, int Hp, int Wp, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DAT-specific (csrc/dat_kernels.cu)
 * ------------------------------------------------------------------------------------------------ */

/* Adaptive Interaction Module tail, dat_arch.py:544-560 (mode 0, spatial block) / :650-664 (mode 1, channel block):
 *   s = w2 . hid[p] + b2, with hid = gelu(W1 . (mode ? conv : att)[p] + b1) computed beforehand by ff_conv_gemm
 *       (spatial_interaction, BN folded into W1/b1; bf16 [M][hid_ld >= 32], columns >= nhid ignored)
 *   mode 0: out = att * cgate[b] + sigmoid(s) * conv;   mode 1: out = att * sigmoid(s) + conv * cgate[b]
 * att / conv / out are bf16 [M][192]; cgate = sigmoid(channel_interaction output) [B][192] fp32. */
int ff_dat_aim(const void* att, int att_ld, const void* conv, int conv_ld, const float* cgate, int cgate_ld, const void* hid,
               int hid_ld, const float* w2, float b2, int nhid, int mode, long long M, int pixels_per_sample, void* out,
               int out_ld, void* stream);

/* Channel attention weights of AdaptiveChannelAttention (dat_arch.py:632-646): per (sample, head)
 * softmax_j( <q_i,k_j> / (max(|q_i|,1e-12) max(|k_j|,1e-12)) * temperature[h] ) with the contraction over all N tokens,
 * emitted as a block-diagonal bf16 [B*192][192] matrix so `attn @ v` runs through ff_conv_gemm (w_batch_rows=192).
 * scratch >= B*heads*ceil(N/512)*1088 floats. */
int ff_dat_channel_attention_weights(const void* qkv, int ld, int q_off, int k_off, int B, int N, int heads, int hd,
                                     const float* temperature, void* wout, float* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fusion head (csrc/freq.cu, csrc/head_kernels.cu)
 * ------------------------------------------------------------------------------------------------ */

/* 9-band frequency decomposition of the LR tile: lr fp32 NCHW [B,3,H,W] -> bands fp32 [B*H*W][27], channel = band*3 + c,
 * bands = [DCT low/mid/high, DWT LL/LH/HL/HH, FFT low/high].  Replaces MultiDomainFrequencyDecomposition.decompose
 * (multi_domain_frequency.py:578-591; DCT :146-196, DWT :251-299, FFT :352-385; torch.fft -> four DFT passes).
 * dct_mat: [64] DCT-II matrix D[k][n]; dct_band_of: [64] band id (0/1/2) of each coefficient; dwt_lo/hi: db4 taps [8];
 * fft_mask: [H][W/2+1] = sigmoid(bilinear(freq_mask_logits) * clamp(temperature, 1)).  scratch >= B*3*(4*H*(W/2+1) + 4*(H/2+4)*(W/2+4)) floats. */
int ff_freq_decompose(const float* lr, int B, int H, int W, const float* dct_mat, const int* dct_band_of, const float* dct_scale,
                      const float* dwt_lo, const float* dwt_hi, const float* dwt_scale, const float* fft_mask, const float* fft_scale,
                      float* bands, float* scratch, size_t scratch_bytes, void* stream);

/* Cross-band attention front end (large_kernel_attention.py:218-226): per (pixel, band) token, band_proj (1x1, 3->64) then
 * LayerNorm(64); writes the un-normalised projection (bf16, residual) and the normalised tokens (bf16, MHA input). */
int ff_cb_embed_ln(const float* bands, long long tokens, const float* proj_w, const float* proj_b, const float* ln_w, const float* ln_b,
                   void* stacked, void* normed, void* stream);
/* Core of nn.MultiheadAttention over the num_bands tokens of each pixel (4 heads x 16, :228): qkv bf16 [tokens][192] -> bf16 [tokens][64]. */
int ff_cb_attention(const void* qkv, long long tokens, int num_bands, void* out, void* stream);
/* Same core for `group` tokens of width dim = 64 (4 heads) or 128 (8 heads): qkv bf16 [tokens][3*dim], out bf16 [tokens][dim].  dim 128 /
 * group 3 is the cross-expert attention of the collaborative branch (large_kernel_attention.py:375-381). */
int ff_token_attention(const void* qkv, long long tokens, int group, int dim, void* out, void* stream);
/* Collaborative branch of forward_with_precomputed(..., expert_features) (large_kernel_attention.py:327-419):
 *  - NCHW fp32 expert features -> NHWC bf16 GEMM operand rows [B*H*W][ld], columns >= C zero (:341-356);
 *  - partial[b][blk][c] = sum over HR pixels of GELU(bilinear_up_scale(g)[.][c]), g fp32 [B*h*w][ld] = the modulation head's first 1x1
 *    conv at LR resolution (it commutes with the up-sampling); ff_gap_finalize(partial, B, nblk, C, 1/(HR pixels)) ends the
 *    AdaptiveAvgPool2d(1) of :407-411;
 *  - x[p][c_off + c] = clamp(x * (f0 + f1 * m[b][c]), 0, 1) on fp32 rows: out * (1 + 0.2 (mod - 0.5)), clamped (:414-415). */
int ff_nchw_to_nhwc_bf16(const float* x, int B, int C, int H, int W, void* out, int ld, void* stream);
int ff_up_gelu_pool(const float* g, int ld, int B, int h, int w, int C, int scale, int nblk, float* partial, void* stream);
int ff_scale_clamp_channels(float* x, int ld, int B, long long pixels_per_sample, int c_off, int C, const float* m, int m_ld, float f0, float f1,
                            void* stream);
/* y = x * a[c] + b[c] on bf16 rows (eval-mode BatchNorm that cannot be folded through a zero-padded depthwise conv, :145). */
int ff_affine_rows(const void* x, long long rows, int C, const float* a, const float* b, void* y, void* stream);
/* AdaptiveBandFusionModule 9->3 (multi_domain_frequency.py:478-526) fused with the frequency guidance of
 * enhanced_fusion.py:533-542.  blob layout: see csrc/head_kernels.cu (band_fuse_kernel). */
int ff_band_fuse(const float* bands, const float* att, int att_ld, long long P, const float* blob, int blob_len, float* feats,
                 float* guidance, void* stream);
/* fp32 NHWC bilinear resize, align_corners=False (F.interpolate call sites of fusion_network.py:594-603); accumulate!=0 adds into out. */
int ff_bilinear_f32(const float* in, int B, int Hi, int Wi, int ld_in, int C, float* out, int Ho, int Wo, int ld_out, int accumulate,
                    const float* bias, void* stream);
/* Same with the source-coordinate ratio given instead of derived from the sizes: F.interpolate(scale_factor=s) maps
 * src = (dst + 0.5) / s - 0.5 while the output size is floor(in * s) -- the two differ on sizes s does not divide (fusion_network.py:594,599). */
int ff_bilinear_f32_scaled(const float* in, int B, int Hi, int Wi, int ld_in, int C, float* out, int Ho, int Wo, int ld_out, float ratio_y,
                           float ratio_x, void* stream);
/* bf16 NHWC bilinear x2 (hierarchical_fusion.py:155-158,177-180) written into channels [0,C) of a wider row. */
int ff_bilinear_up2_bf16(const void* in, int B, int Hi, int Wi, int ld_in, int C, void* out, int ld_out, void* stream);
/* DynamicExpertSelector tail (fusion_network.py:221-234), in place on [P][4] = (gate0, gate1, gate2, difficulty). */
int ff_selector_tail(float* gates_difficulty, long long P, void* stream);
/* Bilinear resize (factor 1, 2 or 4 down) of the 9 stacked expert channels into a bf16 conv-input buffer (hierarchical_fusion.py:140-171).
 * With 16-byte aligned operands (ld % 4 == 0, out_off % 8 == 0, out_ld >= out_off + 16) the channels out_off + 9 .. out_off + 15 are written
 * as zeros (they are zero padding of the conv input rows). */
int ff_experts_resize(const float* stack, int ld, int B, int H, int W, int factor, void* out, int out_ld, int out_off, void* stream);
/* SpatialGate (hierarchical_fusion.py:25-43), in place: x *= sigmoid(w2 . gelu(W1 x + b1) + b2); C in {32, 64}. */
int ff_pixel_gate(void* x, int ld, long long P, int C, const float* w1, const float* b1, const float* w2, float b2, void* stream);
/* HR blend: 0.7*hier + 0.3*freq-weighted experts, dynamic selection (enhanced_fusion.py:550-556, 593-647); also emits
 * base = fused + residual_scale * bilinear_up4(lr) (:677-681) for the refine-net epilogue. */
int ff_blend(const float* stack, int ld_s, const float* hier, const float* guidance, const float* gates, const float* lr, int B, int h, int w,
             float residual_scale, float* fused, float* base, void* stream);
/* Laplacian pyramid (edge_enhancement.py:182-220): down = avg_pool2(gaussian5x5(cur)); lap = cur - bilinear_up2(down). */
int ff_gauss_down(const float* cur, int ld, int B, int H, int W, const float* k1d, float* down, int ld_o, void* stream);
int ff_lap_sub(const float* cur, int ld, const float* down, int ld_d, int B, int H, int W, float* lap, int ld_l, void* stream);
/* out[:, off:off+C] = weight * bilinear_up(feat * att) for one pyramid level (edge_enhancement.py:240-250). */
int ff_edge_merge(const void* feat, int ld_f, const float* att, int B, int Hl, int Wl, int C, float weight, void* out, int H, int W, int ld_o,
                  int off, void* stream);
/* out_nchw = clamp(sr + gate * strength * edge, 0, 1)  (edge_enhancement.py:256-260); se = [P][8] (sr 0..2, edge 3..5). */
int ff_edge_final(const float* se, const float* gate, int gate_ld, int B, int H, int W, float strength, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tile scheduler tail (csrc/stitch.cu)
 * ------------------------------------------------------------------------------------------------ */

/* Overlapped-tile stitch of io._tiled_forward (models/team29_FreqFusion/io.py:97-121): ramp-weighted accumulation in the
 * reference's tile order with un-fused fp32 ops (bit-identical), normalisation by clamp(weight_map, 1e-8), and optionally
 * io._save_image's quantisation (:71-76) round_half_even(clamp(x,0,1)*255) into an HWC uint8 image.
 * tiles: fp32 [ny*nx][3][ts][ts] (y-major, x-minor); ty/tx: HR origin of each tile row / column; wy/wx: [ny][ts] / [nx][ts]
 * 1-D blend weights; out: fp32 [3][H][W] or NULL; out_u8: uint8 [H][W][3] or NULL. */
int ff_stitch(const float* tiles, const int* ty, const int* tx, const float* wy, const float* wx, int ny, int nx, int ts, int H, int W,
              float* out, unsigned char* out_u8, void* stream);

/* io._load_image (io.py:64-68) fused with the tile extraction of _tiled_forward (:99-103): HWC uint8 image on the device ->
 * fp32 NCHW tiles [ny*nx][3][th][tw], value = (float)u8 / 255.0f (IEEE division, bit-identical to the reference's numpy
 * expression).  ys/xs: LR origins of the tile rows / columns (device int32).  ny = nx = 1 with origin 0 converts a whole image. */
int ff_u8_to_tiles(const unsigned char* img, int H, int W, const int* ys, const int* xs, int ny, int nx, int th, int tw, float* tiles,
                   void* stream);

/* io._save_image (io.py:71-76) for un-tiled results: fp32 NCHW [3][H][W] -> uint8 HWC, round_half_even(clamp(x,0,1) * 255). */
int ff_quantize_u8(const float* x, int H, int W, unsigned char* out, void* stream);

/* PSNR (dB) on the BT.601 Y channel with a `crop`-pixel border removed, per sample: a, b fp32 NCHW [B,3,H,W] in [0,1]
 * (reference src/utils/metrics.py:30-52, 76-126).  out: fp32 [B]; scratch >= B*64 doubles.  Identical inputs give 100 dB. */
int ff_psnr_y(const float* a, const float* b, int B, int H, int W, int crop, float* out, double* scratch, size_t scratch_bytes, void* stream);

/* SSIM on the BT.601 Y channel with a `crop`-pixel border removed, per sample, same operands as ff_psnr_y (inputs are clamped
 * to [0,1] first, as the reference does).  Reference src/utils/metrics.py:189-246 in its PyTorch branch :129-186 (taken when
 * scikit-image is absent): 11x11 Gaussian window sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2, mean over the cropped
 * map.  out: fp32 [B]; scratch >= ff_ssim_y_scratch_bytes(B, H, W, crop) bytes (one double per 32x32 map tile and sample). */
int ff_ssim_y(const float* a, const float* b, int B, int H, int W, int crop, float* out, double* scratch, size_t scratch_bytes, void* stream);
size_t ff_ssim_y_scratch_bytes(int B, int H, int W, int crop);

/* The PSNR / SSIM pair of the reference's evaluation harness (eval.py:157 -> utils/utils_image.py:287-312 cal_psnr_ssim) on two
 * uint8 RGB images [H][W][3] in device memory: crop `border` pixels, Y = OpenCV's 8-bit RGB2YCrCb luma
 * ((4899 R + 9617 G + 1868 B + 2^13) >> 14), PSNR = 10 log10(255^2 / MSE) (inf for identical images), SSIM = scikit-image's
 * structural_similarity defaults (7x7 uniform window, sample covariance, K1 0.01, K2 0.03, data_range 255, mean over the windows
 * inside the cropped image); integer window sums, fp64 formula.  out: device double[2] = {psnr, ssim};
 * scratch >= ff_eval_scratch_bytes(H, W, border) bytes, 8-byte aligned. */
int ff_eval_psnr_ssim_u8(const unsigned char* a, const unsigned char* b, int H, int W, int border, double* out, void* scratch,
                         size_t scratch_bytes, void* stream);
size_t ff_eval_scratch_bytes(int H, int W, int border);

#ifdef __cplusplus
}
#endif
#endif
