/*
 * laud_b200.h - C ABI of the B200-native LAUDNet dynamic-operator hot path.
 *
 * One shared library (liblaud_b200.so, sm_100a) exports everything below.
 * All pointers are DEVICE pointers unless marked `host`; activations are
 * fp16 NHWC ("channels last"); masks are uint8 0/1; index lists int32; all
 * reductions that feed a gating decision accumulate in fp32 in a fixed order.
 * Every entry point enqueues work on `stream` (a cudaStream_t passed as
 * void*), performs no host synchronisation and no allocation (workspaces are
 * caller-provided), returns 0 on success or a negative LAUD_E_* code, and
 * leaves a thread-local message for laud_last_error().
 *
 * The reference (LeapLabTHU/LAUDNet) is pure Python and has no FFI; each entry
 * point names the reference interface it replaces (paths relative to the
 * reference checkout, imagenet_classification/models/...).  The Python
 * binding a maintainer would add is shown in INTEGRATION.md.
 */
#ifndef LAUD_B200_H_
#define LAUD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAUD_ABI_VERSION 4

enum {
  LAUD_OK = 0,
  LAUD_E_BADARG = -1,      /* shape / alignment / unsupported combination */
  LAUD_E_CUDA = -2,        /* a CUDA runtime call failed (message has the cudaError string) */
  LAUD_E_UNSUPPORTED = -3  /* valid in the reference, not implemented here (raises, never falls back) */
};

/* which implementation of the gather-GEMM convolution to run */
enum {
  LAUD_CONV_AUTO = 0,   /* product path: tcgen05/TMEM kernel */
  LAUD_CONV_UMMA = 1,   /* tcgen05.mma + TMEM accumulators, swizzled smem staging */
  LAUD_CONV_HMMA = 2,   /* legacy warp-level tensor-core kernel (cross-check / bring-up) */
  LAUD_CONV_NAIVE = 3   /* one thread per output, fp32 (self-test only) */
};

/* relu_mode of laud_conv_desc */
enum {
  LAUD_RELU_NONE = 0,
  LAUD_RELU_ALL = 1,
  LAUD_RELU_WHERE_GATE0 = 2 /* relu only where out_mask == 0 (downsample branch of a spatially skipped block) */
};

#define LAUD_GAP_SPLITS 8        /* row splits of the deterministic two-phase global average pool */
#define LAUD_PREBIAS_CLASSES 16  /* border classes of a 3x3/pad-1 conv: 4 row patterns x 4 col patterns */

int laud_abi_version(void);
const char* laud_last_error(void);
/* number of kernels this library has launched in this process (all streams) */
unsigned long long laud_launch_count(void);
/* launches of the mask-conditioned convolution by implementation:
 * out[0] tcgen05/TMEM kernel, out[1] legacy HMMA kernel, out[2] naive self-test kernel */
void laud_conv_path_counts(unsigned long long out[3]);
/* launches of the TMA-staged tcgen05 kernel (conv_tma.cu), a subset of out[0] above */
unsigned long long laud_conv_tma_launch_count(void);
/* Measurement aid: while enabled, every tcgen05 conv launch is bracketed - the kernel only, not the host-side
 * descriptor encoding - by CUDA events on its stream.  laud_conv_profile_read (after a synchronize) returns the
 * number of recorded launches and their summed device time. */
void laud_conv_profile(int enable);   /* 1 start (drops old records), 2 stop but keep the records, 0 stop and drop */
int laud_conv_profile_read(float* total_ms);
/* per-launch device times in launch order.  When the launches were CAPTURED into a CUDA graph while profiling was on,
 * the event records are nodes of that graph: call this after a replay (and a synchronize) to read the times of the
 * kernels inside that graph execution. */
int laud_conv_profile_read_all(float* ms_out, int cap);

/* ---------------------------------------------------------------------------
 * (a1) channel masker.  Replaces Masker_channel_MLP.forward, eval branch
 * (models/utils.py:113-131): GAP -> [Linear -> ReLU ->] Linear -> keep>=drop.
 *   x        fp16 [B, HW, C] (C % 8 == 0)
 *   layers   2: w1[hidden,C] b1[hidden] w2[2G,hidden] b2[2G]   (fp32)
 *            1: w1[2G,C]     b1[2G]     (w2,b2 NULL, hidden ignored)
 *   partial_ws  fp32 [B, LAUD_GAP_SPLITS, C] scratch
 *   pooled_out  fp32 [B, C]      (nullable)  the GAP values
 *   logits_out  fp32 [B, 2G]     (nullable)  first G = keep, last G = drop
 *   mask_out    u8   [B, G]
 *   idx_out     i32  [B, G]: ascending ACTIVE group ids in [0,cnt), then the
 *                             ascending INACTIVE ids in [cnt,G)
 *   cnt_out     i32  [B]
 *   total_out   i32  [1]   += sum_b cnt[b]  (caller zeroes it; feeds rho_c)
 * ------------------------------------------------------------------------- */
int laud_masker_channel_mlp(const void* x, int B, int HW, int C, int layers,
                            const float* w1, const float* b1, int hidden,
                            const float* w2, const float* b2, int G,
                            float* partial_ws, float* pooled_out, float* logits_out,
                            uint8_t* mask_out, int32_t* idx_out, int32_t* cnt_out,
                            int32_t* total_out, void* stream);

/* Decision + compaction only, from caller-provided pooled features [B,C]
 * (used by Masker_channel_conv_linear, models/utils.py:150-169, whose pooled
 * input is the GAP of a conv-BN-ReLU branch). Same outputs as above. */
int laud_masker_channel_from_pooled(const float* pooled, int B, int C, int layers,
                                    const float* w1, const float* b1, int hidden,
                                    const float* w2, const float* b2, int G,
                                    float* logits_out, uint8_t* mask_out, int32_t* idx_out,
                                    int32_t* cnt_out, int32_t* total_out, void* stream);

/* Same decision from the fused-GAP partial sums a producing laud_conv_forward left in
 * laud_conv_desc::gap_partial (fp32 [B, gap_tiles, C], layout documented there): pooled[b,c] =
 * (sum of sample b's partials in ascending k) / HW, then MLP -> keep>=drop -> compaction as
 * laud_masker_channel_mlp.  Replaces the GAP read of Masker_channel_MLP.forward (models/utils.py:113-117). */
int laud_masker_channel_from_partials(const float* partials, int B, int HW, int C, int gap_tiles, int layers,
                                      const float* w1, const float* b1, int hidden,
                                      const float* w2, const float* b2, int G,
                                      float* pooled_out, float* logits_out, uint8_t* mask_out,
                                      int32_t* idx_out, int32_t* cnt_out, int32_t* total_out, void* stream);

/* Deterministic global average pool of fp16 [B,HW,ldx] (first C channels) -> fp32 [B,C]. */
int laud_global_avg_pool(const void* x, int B, int HW, int C, int ldx,
                         float* partial_ws, float* pooled_out, void* stream);

/* ---------------------------------------------------------------------------
 * (a2) spatial / layer masker.  Replaces Masker_spatial.forward, eval branch
 * (models/utils.py:47-65): adaptive_avg_pool2d(x, S) [skipped if S >= H]
 * -> 1x1 conv C->2g (+bias) -> keep>=drop.   S == 1 is the layer gate.
 *   x fp16 [B,H,W,C];  w fp32 [2g,C];  bias fp32 [2g]
 *   logits_out fp32 [B,2g,S,S] (nullable);  mask_out u8 [B,g,S,S]
 *   total_out  i32 [1] += number of ones
 * ------------------------------------------------------------------------- */
int laud_masker_spatial(const void* x, int B, int H, int W, int C,
                        const float* w, const float* bias, int g, int S,
                        float* logits_out, uint8_t* mask_out, int32_t* total_out, void* stream);

/* (a4) ExpandMask.forward (models/utils.py:74-89): zero-insert upsample by
 * `stride`, then OR over a (2*padding+1)^2 window and over all g groups.
 *   mask u8 [B,g,H,W] -> out u8 [B,g,H*stride,W*stride]; total_out += ones. */
int laud_expand_mask(const uint8_t* mask, int B, int g, int H, int W, int stride, int padding,
                     uint8_t* out, int32_t* total_out, void* stream);

/* F.interpolate(mode='nearest') of a mask (laud_resnet.py:106): src = floor(dst*S/H). */
int laud_resize_mask_nearest(const uint8_t* mask, int B, int g, int S, int H_out,
                             uint8_t* out, void* stream);

/* The S == 1 case of that resize on a NON-square feature map (the detection backbone resizes the gate to the actual
 * feature size, mmdet/models/backbones/lad_mmdet_resnet.py:274): out[b, j, 0:hw] = gate[b, j] for every (sample, group);
 * total_out += ones written (nullable). */
int laud_broadcast_gate(const uint8_t* gate /* [B, g] */, int B, int g, int hw, uint8_t* out /* [B, g, hw] */,
                        int32_t* total_out, void* stream);

/* The three spatial masks of a block in one launch (laud_resnet.py:105-110 = laud_resize_mask_nearest +
 * laud_expand_mask(1,0) + laud_expand_mask(stride,1)): small u8 [B,g,S,S] -> m3, m2 u8 [B,g,H_out,H_out],
 * m1 u8 [B,g,H_out*stride,H_out*stride]; total2 / total1 += ones of m2 / m1 (x g, as the reference's means count). */
int laud_spatial_masks(const uint8_t* small, int B, int g, int S, int H_out, int stride, uint8_t* m3, uint8_t* m2,
                       uint8_t* m1, int32_t* total2, int32_t* total1, void* stream);

/* Ordered compaction of a gate: rows_out[0..n) = ascending flat indices i with
 * gate[i] != 0 (any group), n -> count_out[0].  gate u8 [N] (g==1) or [B,g,HW]
 * (a row is active if any of its g groups is).  Deterministic (single pass,
 * decoupled look-back free: one CTA per 2048 items + ordered prefix). */
int laud_compact_rows(const uint8_t* gate, int B, int g, int HW,
                      int32_t* rows_out, int32_t* count_out, int32_t* block_ws, void* stream);

/* Layer gate (dyn_mode='layer', laud_resnet.py:72,97-110: one gate per sample) in one launch: the ordered list of
 * the ACTIVE samples (the work list of laud_conv_forward's sample_idx / sample_cnt) and the statistics counts of the
 * broadcast / dilated masks: counts4[2] += n_active*hw_out, counts4[3] += n_active*hw_in (counts4 nullable). */
int laud_layer_gate_lists(const uint8_t* gate /* [B] */, int B, int hw_out, int hw_in, int32_t* counts4,
                          int32_t* rows_out /* [B] */, int32_t* count_out /* [1] */, void* stream);

/* (f4) Training-mode gate with SUPPLIED Gumbel noise.  Replaces
 *   F.gumbel_softmax(logits.view(b,2,...), dim=1, tau=temperature, hard=True)[:, 0]
 * of Masker_spatial / Masker_channel_MLP / Masker_channel_conv_linear.forward in training mode (models/utils.py:56-58,
 * 123-125, 161-163) for a noise tensor given by the caller (torch's Philox stream cannot be reproduced in a kernel):
 *   keep <=> (l_keep + g_keep) / tau >= (l_drop + g_drop) / tau        (ties keep, as argmax does)
 * logits / noise fp32 [B, 2, G, inner] (the logits the maskers emit through their logits_out argument; noise may be
 * NULL = the eval decision); mask_out u8 [B, G, inner]; for channel gates (inner == 1) idx_out [B, G] / cnt_out [B] in
 * the maskers' layout (active ids ascending, then inactive); total_out += number of kept decisions. */
int laud_gate_from_logits(const float* logits, const float* noise, int B, int G, int inner, float tau, uint8_t* mask_out,
                          int32_t* idx_out, int32_t* cnt_out, int32_t* total_out, void* stream);

/* ---------------------------------------------------------------------------
 * (a5-a7) the mask-conditioned convolution: implicit-GEMM conv (1x1 or 3x3)
 * + folded BatchNorm + mask + residual + ReLU, with per-sample channel
 * gathers and row (pixel / sample) gathers.  Replaces the
 * conv -> apply_channel_mask -> bn -> relu / conv -> bn -> apply_spatial_mask
 * -> += identity -> relu chains of Bottleneck.forward (laud_resnet.py:115-144).
 *
 * GEMM view: D[m, n] = sum_k A[m, k] * Wt[n, k]
 *   m = output pixel (b, oy, ox), n = output channel, k = (tap, in-channel).
 * Channel gathers (per sample b, granularity `gran` consecutive channels):
 *   k side: x holds only the ACTIVE input channels of sample b, compacted to
 *           the front (x[b,p,j], j < k_cnt[b]*k_gran); compact channel j is
 *           weight in-channel k_idx[b, j/gran]*gran + j%gran.
 *   n side: y receives only the ACTIVE output channels, compacted the same way.
 * Row gathers: sample_idx/sample_cnt restrict the batch to a device-side list
 *   of samples (layer skip); row_idx/row_cnt restrict to a device-side list of
 *   output pixels, flat index b*H_out*W_out + oy*W_out + ox (spatial skip).
 *   Rows not listed are NOT written (in-place residual semantics).
 * Epilogue, for real output channel o (compact j):
 *   v = acc + pre_bias[b, cls(oy,ox), j]          (H1 constants; optional)
 *   v = v * scale[o] + shift[o]                   (folded BN; optional)
 *   v = v * out_mask[b, o / (C_out/mask_groups), oy, ox]   (optional, u8)
 *   v = v + residual[b, oy, ox, o]                (optional, fp16)
 *   v = relu(v) per relu_mode;  y = fp16(v)
 *   compact channels [n_count, round_up(n_count, n_pad_align)) := 0  (n_pad_align > 0);
 *   with n_idx the rest of the row, [round_up(n_count, n_pad_align), ldy), is SCRATCH: the
 *   TMA-staged kernel stores whole 64-channel slabs, consumers read only the padded width.
 * ------------------------------------------------------------------------- */
typedef struct laud_conv_desc {
  const void* x;  int32_t ldx;          /* fp16 [B,H_in,W_in,ldx] */
  const void* w;                        /* fp16 [C_out, ksize*ksize, C_in] */
  void* y;        int32_t ldy;          /* fp16 [B,H_out,W_out,ldy] */
  int32_t B, H_in, W_in, C_in, H_out, W_out, C_out;
  int32_t ksize, stride, pad;
  const float* scale; const float* shift;      /* [C_out] */
  int32_t relu_mode;
  const void* residual; int32_t ldr;           /* fp16 [B,H_out,W_out,ldr] */
  const int32_t* k_idx; const int32_t* k_cnt; int32_t k_ld; int32_t k_gran;
  const int32_t* n_idx; const int32_t* n_cnt; int32_t n_ld; int32_t n_gran;
  const float* pre_bias; int32_t pre_bias_classes; int32_t pre_bias_ld; /* fp32 [B,classes,pre_bias_ld] */
  const uint8_t* out_mask; int32_t mask_groups;  /* u8 [B,mask_groups,H_out,W_out] */
  const int32_t* sample_idx; const int32_t* sample_cnt;
  const int32_t* row_idx; const int32_t* row_cnt;
  int32_t n_pad_align;      /* 0 | 8 | 16: zero-pad each sample's compact output channels to this multiple */
  float* gap_partial;       /* optional fused global-average-pool of the OUTPUT (feeds the next block's channel masker,
                               models/utils.py:113-117, without re-reading the activations): fp32
                               [B, gap_tiles, C_out] PARTIAL SUMS.  The B*H*W output pixels are cut into tiles of 128
                               consecutive pixels; slot k of sample b holds the sum over the pixels of b inside the
                               k-th tile that contains any of them (k = tile - (b*H*W)/128; slots past the last
                               such tile are not written).  Needs gap_tiles >= (H*W-1)/128 + 2.  Only 1x1 stride-1
                               layers with nothing per sample (no lists, gates, n_mask), C_out % 64 == 0, H*W >= 43;
                               anything else returns LAUD_E_UNSUPPORTED.  laud_masker_channel_from_partials consumes it.
                               Deterministic (no atomics); the fp32 grouping of a sample's sum follows the tile
                               boundaries, i.e. depends on the sample's offset in the batch. */
  int32_t gap_tiles;
  const void* w_t;          /* optional transposed copy of w: fp16 [ksize*ksize, C_in, C_out].  With k_idx it
                               selects the K-row-gather path (16-byte gathers of the active input channels;
                               all output channels are computed and, with n_idx, compacted in the epilogue).
                               In that path pre_bias is indexed by REAL output channel, not compact column. */
  const void* bias_t;       /* optional, w_t path only: H1 constants fp16 [B, ksize*ksize, C_out] (sample pitch
                               bias_ld elements).  Enters the GEMM as one extra K=16 step:
                               acc[p,o] += sum_{taps of p that fall inside the input} bias_t[b,tap,o].
                               Replaces pre_bias (mutually exclusive). */
  int32_t bias_ld;
  const uint8_t* n_mask;    /* optional MASKED-DENSE execution of the channel gate (reference laud_resnet.py:116,124:
                               conv -> x mask -> bn): u8 [B, C_out / n_mask_gran].  Every output channel is computed
                               with weights shared by all samples (no gather); a channel whose gate is 0 is emitted as
                               if its accumulator were 0, i.e. the BN constant shift[o] (then ReLU) - the same values
                               the gathered path reproduces through the H1 constants.  Exclusive with n_idx / k_idx. */
  int32_t n_mask_gran;
  int32_t n_expand;         /* with n_idx (and no k_idx): CHANNEL SKIPPING WITH A DENSE RESULT.  Only the sample's ACTIVE
                               output channels are computed - their weight rows are gathered by the TMA unit
                               (cp.async.bulk.tensor ... tile::gather4 over w, four rows per instruction, indices from
                               n_idx) so the tcgen05 MMAs run over N = 2*n_cnt[b] columns instead of C_out - and the
                               epilogue EXPANDS them to their real positions of a dense row y[b,p,0:C_out]; a gated
                               channel is written as its BN constant shift[o] (then ReLU), exactly what n_mask
                               produces with every MMA executed (reference laud_resnet.py:123-126: conv -> x mask ->
                               bn -> relu).  The consumer therefore reads ordinary dense activations and keeps shared
                               weights.  Needs: 3x3, stride 1, pad 1, W_out + 2 <= 128, even n_gran, scale/shift,
                               relu_mode NONE | ALL, no residual / out_mask / lists; n_idx rows hold ALL group ids
                               (active ascending, then inactive - the layout the maskers emit), n_ld == C_out/n_gran.
                               Anything else returns LAUD_E_UNSUPPORTED. */
} laud_conv_desc;

int laud_conv_forward(const laud_conv_desc* desc /* host */, int impl, void* stream);
/* Launch policy of the TMA-staged convolution kernel: 1 = programmatic dependent launch (the prologue of a kernel - barrier
 * initialisation, TMEM allocation, descriptor prefetch - overlaps the tail of its predecessor in the stream).  Pays for
 * single-chain forwards of small batches (configs[0], batch 8: +4 %), costs with two parallel graph chains; default 0. */
void laud_conv_set_pdl(int enable);

/* H1 constants of channel-skipping with mask-before-BN (laud_resnet.py:115-118,
 * 123-126): a masked channel k of conv1's (conv2's) output is the constant
 * c_k = relu(shift_k) after BN+ReLU, so the exact sparse conv2 / conv3 add
 *   T2[b,tap,o] = sum_{k masked} relu(shift1[k]) * w2[o,tap,k]   (per valid tap)
 *   T3[b,o]     = sum_{k masked} relu(shift2[k]) * w3[o,k]
 * Both are ONE dense GEMM  inact[B,width] x cw[9*width + C_out, width]^T  run with
 * laud_conv_forward (1x1, B=1, H_out*W_out = batch), where cw are the weights
 * pre-scaled by relu(shift) (packed once per model) and inact is the 0/1
 * indicator of the masked channels:
 *   laud_gate_inactive: mask u8 [B,G] -> inact fp16 [B, G*gran] (1.0 where masked)
 * On the w_t path T feeds the convolutions directly (laud_conv_desc.bias_t: one extra K=16
 * MMA step, no pre_bias).  For the other layouts:
 *   laud_channel_consts_fold: T fp16 [B, 9*width + C_out] (the GEMM output, T2 rows tap-major) ->
 *     pre_bias2 fp32 [B,16,width]: the 9 taps folded into the 16 border classes,
 *       indexed by compact active output channel (compact_index=1, needs idx/cnt)
 *       or by real output channel (compact_index=0, the w_t path);
 *     pre_bias3 fp32 [B,C_out]. */
int laud_gate_inactive(const uint8_t* mask, int B, int G, int gran, void* inact_f16, void* stream);
int laud_channel_consts_fold(const void* T, int B, int width, int C_out, const int32_t* idx,
                             const int32_t* cnt, int G, int gran, int compact_index,
                             float* pre_bias2 /* [B,16,width] */, float* pre_bias3 /* [B,C_out] */,
                             void* stream);

/* ---------------------------------------------------------------------------
 * (a8) network ends.  Stem: conv7x7/2 + BN + ReLU + maxpool3x3/2 fused
 * (laud_resnet.py:317-324).  x fp16 NCHW [B,3,H,W] -> y fp16 NHWC [B,H/4,W/4,C0].
 *   w fp16 [C0,3,7,7] (reference layout), scale/shift fp32 [C0].
 * Head: global avgpool + fc (laud_resnet.py:349-356).
 *   x fp16 [B,HW,C] -> logits fp32 [B,n_cls]; w fp16 [n_cls,C], bias fp32.
 * ------------------------------------------------------------------------- */
int laud_stem_forward(const void* x_nchw, int B, int H, int W, const void* w, int C0,
                      const float* scale, const float* shift, void* y_nhwc, void* stream);
int laud_head_forward(const void* x, int B, int HW, int C, const void* w, const float* bias,
                      int n_cls, float* pooled_ws /* [B,C] */, float* logits, void* stream);
/* The head from the fused-GAP partial sums the last convolution left (laud_conv_desc::gap_partial,
 * fp32 [B, gap_tiles, C]): pooled_ws fp32 [B, C] receives the pooled features, then the same fc. */
int laud_head_forward_from_partials(const float* partials, int B, int HW, int C, int gap_tiles,
                                    const void* w, const float* bias, int n_cls,
                                    float* pooled_ws, float* logits, void* stream);

/* Layout helpers (operator-level API / tests): fp32|fp16 NCHW <-> fp16 NHWC. */
int laud_nchw_to_nhwc_f16(const void* src, int src_is_f32, int B, int C, int H, int W,
                          void* dst, int ldd, void* stream);
int laud_nhwc_f16_to_nchw_f32(const void* src, int lds, int B, int C, int H, int W,
                              float* dst, void* stream);

/* ---------------------------------------------------------------------------
 * (a9) LAUD-RegNet-Y operators (models/laud_regnet.py).  The 1x1 convolutions a / c / proj, the maskers and the head
 * run on the entry points above.
 *   laud_regnet_stem_forward     SimpleStemIN conv3x3/2 + BN + ReLU (:59-71): x fp16 NCHW [B,3,H,W] -> y fp16 NHWC
 *                                [B,H/2,W/2,C0]; w fp16 [C0,3,3,3]
 *   laud_grouped_conv3x3_forward the transform's conv b (:118-120,188): grouped 3x3 (pad 1, stride 1|2) + BN + ReLU,
 *                                w fp16 [C][9][group_width] (out channel, tap, in channel of its group); group widths 8 / 16 / 24
 *                                on dedicated kernels with an optional channel gate ch_mask u8 [B, C/mask_gran] applied to the
 *                                input (= conv a's gated output, :183) and to the output (:189); any wider multiple of 8
 *                                (RegNetY-8GF / 16GF / 32GF) runs group by group on the tcgen05 convolution kernel (no gate)
 *   laud_se_gate                 SqueezeExcitation gate (:128-132,194) from the pooled features fp32 [B,C]:
 *                                gate = sigmoid(W2 relu(W1 p + b1) + b2)  (p and gate multiplied by ch_mask if given);
 *                                w1 fp32 [S][C], w2 fp32 passed TRANSPOSED as [S][C]
 *   laud_scale_channels          x[b,p,c] *= gate[b,c]  (fp16 NHWC, in place)
 * ------------------------------------------------------------------------- */
int laud_regnet_stem_forward(const void* x_nchw, int B, int H, int W, const void* w, int C0, const float* scale,
                             const float* shift, void* y_nhwc, void* stream);
int laud_grouped_conv3x3_forward(const void* x, int B, int H_in, int W_in, int C, int stride, const void* w,
                                 int group_width, const float* scale, const float* shift, const uint8_t* ch_mask,
                                 int mask_gran, void* y, void* stream);
int laud_se_gate(const float* pooled, int B, int C, const float* w1, const float* b1, int S, const float* w2,
                 const float* b2, const uint8_t* ch_mask, int mask_gran, float* gate, void* stream);
int laud_scale_channels(void* x, int B, int HW, int C, const float* gate, void* stream);

/* ---------------------------------------------------------------------------
 * Forward statistics.  Reproduces, in one launch and in the reference's fp32
 * evaluation order, the per-block densities / flops_perc / flops that
 * Bottleneck.forward and ResNet.forward thread through the network
 * (laud_resnet.py:112-162, 321-356).
 *   counts i32 [n_blocks,4]: ones in (channel mask, mask_conv3 small, mask_conv2, mask_conv1)
 *   consts i64 [n_blocks,12]: masker MACs (channel, spatial), c1, c2, c3, projection MACs, the four density
 *          denominators, flags (1 channel gate, 2 spatial gate, 4 RegNet accumulation order), SE flops
 *   out    f32 [n_blocks,5 + 1]: rho3,rho2,rho1,rho_c,flops_perc per block, then total flops
 * ------------------------------------------------------------------------- */
int laud_forward_stats(const int32_t* counts, const int64_t* consts, int n_blocks,
                       int64_t stem_flops, int64_t pool_flops, int64_t fc_flops, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAUD_B200_H_ */
