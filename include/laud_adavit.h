/*
 * laud_adavit.h - C ABI of the AdaViT token / head / layer-skip block (BASELINE.json configs[3]), exported by the same
 * liblaud_b200.so as laud_b200.h (same conventions: device pointers, a cudaStream_t passed as void*, no host
 * synchronisation, no allocation, 0 or a negative LAUD_E_* code + laud_last_error()).
 *
 * PARITY UNPINNED / self-oracle: the reference tree contains no AdaViT model code (README.md:24-26 links the external
 * repository); the only in-tree description is the operator list of DyNetSimulator/adavit/simulate_adavit.py:83-182.
 * Each entry point names the operator of that list it executes; the arithmetic it must reproduce is the declared
 * self-oracle oracle/adavit_oracle.py.
 *
 * Execution scheme (one transformer block):
 *   laud_adavit_policy   decisions of the block from the fp32 residual stream x [B, L, D]
 *   laud_adavit_lists    per-sample row offsets of the two COMPACT token lists (attention / MLP sub-layer)
 *   laud_adavit_ln_gather LayerNorm of the kept tokens only, written as consecutive fp16 rows (+ their destinations)
 *   laud_adavit_row_lists / laud_adavit_ln_rows  the same in two steps: row lists of both sub-layers, LayerNorm over a row list
 *   laud_tok_gemm        tcgen05 GEMM over the compact rows (device-side row count): QKV, proj, fc1 (+GELU), fc2;
 *                        proj / fc2 add their result straight into the residual stream at the rows' destinations
 *   laud_adavit_attention softmax(QK^T)V over the kept tokens of each (sample, kept head)
 * Dropped tokens, heads, sub-layers and samples cost no FLOPs and no bytes beyond the policy pass.
 */
#ifndef LAUD_ADAVIT_H_
#define LAUD_ADAVIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { LAUD_ACT_NONE = 0, LAUD_ACT_GELU = 1 /* exact: 0.5 x (1 + erf(x / sqrt 2)) */ };

/* ---------------------------------------------------------------------------
 * Token GEMM ("dylinear" / "linear" of simulate_adavit.py:87-97,130-147,150-165):
 *   R[i, n] = act( sum_k A[i, k] * W[n, k] + bias[n] ),  i < *row_cnt
 *   out  != NULL: out[i, n]  = fp16(R[i, n])                       (compact rows stay compact)
 *   resid!= NULL: resid[row_idx[i], n] += R[i, n]   (fp32)          (residual add at the token's own row: the
 *                 "x = x + sublayer(x)" of a kept token; every (row, n) is owned by one thread - no atomics)
 * A fp16 [rows_max, lda], W fp16 [N, K] (nn.Linear layout), K % 64 == 0, N % 8 == 0, lda / ldo % 8 == 0, ldres % 4 == 0.
 * row_cnt (device, nullable = rows_max) bounds the rows that are computed: m-tiles past it are never scheduled.
 * Head skipping on the output columns (the QKV "dylinear" with oc_density = head density): with col_gate, n-tile t
 * (columns [t*bn, (t+1)*bn)) of an m-tile is computed only if col_gate[s * gate_ld + t] != 0 for at least one sample s
 * among the samples of the tile's rows (row_sample[i] = sample of compact row i, ascending); skipped tiles are not written.
 * bn: n-tile width, one of 64 / 128 / 192 / 256 (0 = chosen by the library).
 * For K <= 384 the CTA keeps the whole [bn, K] weight tile of ONE n-tile resident in shared memory and streams activation
 * tiles only; longer reductions stream both operands.
 * ------------------------------------------------------------------------- */
typedef struct laud_tok_gemm_desc {
  const void* a; int32_t lda;
  const void* w;
  const float* bias;
  int32_t rows_max, K, N;
  const int32_t* row_cnt;
  int32_t act;
  void* out; int32_t ldo;
  float* resid; int32_t ldres;
  const int32_t* row_idx;
  const uint8_t* col_gate; int32_t gate_ld;
  const int32_t* row_sample;
  int32_t bn;
  int32_t cta_pair;   /* 1: run in clusters of two CTAs (tcgen05.mma.cta_group::2: M = 256 across the pair, each CTA stages its own
                         128 rows and HALF of every weight tile; commits multicast to both CTAs) where the shape allows (no col_gate,
                         bn % 32 == 0).  Same results; measured slower on configs[3], hence opt-in.  0: single CTAs. */
} laud_tok_gemm_desc;

int laud_tok_gemm(const laud_tok_gemm_desc* desc /* host */, void* stream);
/* launches of the tcgen05 token GEMM in this process */
unsigned long long laud_tok_gemm_launch_count(void);

/* ---------------------------------------------------------------------------
 * The MLP sub-layer in ONE kernel (simulate_adavit.py:130-147: fc1 "dylinear", GELU, fc2) with the hidden activations
 * kept on chip:   resid[row_idx[i], 0:D] += GELU(y[i] W1^T + b1) W2^T + b2     for the compact rows i < *row_cnt.
 * Per 128-row tile the normalised rows stay in shared memory; the hidden dimension is walked in chunks of 128 - GEMM1 into
 * TMEM, + b1, GELU, fp16 into a swizzled shared-memory tile that is the A operand of GEMM2 (accumulated in TMEM over all
 * chunks) - and the result is reduced into the residual stream.  The [rows, Hd] hidden tensor (166 MB per DeiT-S layer at
 * batch 512) that laud_tok_gemm(fc1) + laud_tok_gemm(fc2) write and re-read never exists.
 *   y fp16 [rows_max, D]; w1 fp16 [Hd, D], b1 fp32 [Hd]; w2 fp16 [D, Hd], b2 fp32 [D]; D = 128 | 256 | 384, Hd % 128 == 0.
 * ------------------------------------------------------------------------- */
int laud_adavit_mlp_fused(const void* y, int rows_max, int D, int Hd, const int32_t* row_cnt, const void* w1, const float* b1,
                          const void* w2, const float* b2, float* resid, int ldres, const int32_t* row_idx, void* stream);

/* Patch embedding, first half (the unfold of simulate_adavit.py:61): x fp16 NCHW [B, 3, S, S] -> patches fp16
 * [B * (S/P)^2, 3*P*P], column = c*P*P + iy*P + ix (the flattened conv weight's order); P % 8 == 0.  The projection
 * itself is laud_tok_gemm with resid = the token stream and row_idx[i] = b*L + 1 + p. */
int laud_vit_patchify(const void* x_nchw, int B, int S, int P, void* patches, void* stream);
/* Token stream initialisation: x[b, l, :] = pos[l, :] + (l == 0 ? cls : 0); x fp32 [B, L, D], pos fp32 [L, D], cls fp32 [D]. */
int laud_vit_init_tokens(float* x, int B, int L, int D, const float* pos, const float* cls, void* stream);

/* ---------------------------------------------------------------------------
 * Decisions of one block (simulate_adavit.py:150-160 layer / head policy, :99-106 token score), eval mode:
 *   p = LN(x[b,0]; np_w, np_b);  layer_logits[b, 0:2] = ls_w p + ls_b;  head_logits[b, 0:H] = hs_w p + hs_b
 *   tok_logits[b, l] = ts_w . LN(x[b,l]; n1_w, n1_b) + ts_b   (l >= 1; the class token is always kept)
 *   decision = logit >= 0.  A NULL weight pointer switches that policy off (all kept).
 * x fp32 [B, L, D] (D % 32 == 0, D <= 1024, L <= 1024); LayerNorm eps given; fp32 reductions in a fixed order.
 * Outputs: tok_mask u8 [B, L], tok_cnt i32 [B], head_sel u8 [B, H], layer_sel u8 [B, 2]; logits nullable
 * (tok_logits fp32 [B, L] with [b, 0] = 0, head_logits fp32 [B, H], layer_logits fp32 [B, 2]).
 * ------------------------------------------------------------------------- */
int laud_adavit_policy(const float* x, int B, int L, int D, int H, float eps,
                       const float* n1_w, const float* n1_b, const float* ts_w, const float* ts_b,
                       const float* np_w, const float* np_b, const float* ls_w, const float* ls_b,
                       const float* hs_w, const float* hs_b,
                       uint8_t* tok_mask, int32_t* tok_cnt, uint8_t* head_sel, uint8_t* layer_sel,
                       float* tok_logits, float* head_logits, float* layer_logits, void* stream);

/* Row offsets of the two compact token lists of a block: off_attn[b] = sum_{b' < b} layer_sel[b',0] * tok_cnt[b']
 * (off_attn[B] = total rows = the row_cnt of the attention-side GEMMs), off_mlp likewise with layer_sel[b',1].
 * i32 [B + 1] each.  One CTA, ordered (deterministic). */
int laud_adavit_lists(const int32_t* tok_cnt, const uint8_t* layer_sel, int B, int32_t* off_attn, int32_t* off_mlp,
                      void* stream);

/* LayerNorm + gather: for every sample b with off[b+1] > off[b], its kept tokens (tok_mask, ascending l; NULL = all L)
 * are normalised (weight / bias, eps) and written as fp16 rows y[off[b] + rank, 0:D]; row_idx[off[b] + rank] = b*L + l,
 * row_sample[...] = b (both nullable).  x fp32 [B, L, D].  (layernorm of simulate_adavit.py:171,177 on kept tokens only) */
int laud_adavit_ln_gather(const float* x, int B, int L, int D, float eps, const float* w, const float* bias,
                          const uint8_t* tok_mask, const int32_t* off, void* y, int32_t* row_idx, int32_t* row_sample,
                          void* stream);

/* The same job in two steps with an evenly loaded LayerNorm (what AdaViT.forward uses; results bit-identical to
 * laud_adavit_ln_gather).  laud_adavit_row_lists: for every sample, the destination rows b*L + l of its kept tokens
 * (tok_mask, ascending l; NULL = all L) at rows_attn[off_attn[b] + rank] (+ samp_attn[...] = b, nullable) when the sample
 * runs its attention sub-layer, and at rows_mlp[off_mlp[b] + rank] when it runs its MLP; either (off, rows) pair may be
 * NULL.  laud_adavit_ln_rows: y[r, 0:D] = LayerNorm(x[row_idx[r], 0:D]) as fp16 for r < min(*row_cnt, rows_max)
 * (row_cnt NULL = rows_max); x fp32 rows of D.  (L_select tokens :108, layernorm :171,177 of simulate_adavit.py) */
int laud_adavit_row_lists(const uint8_t* tok_mask, int B, int L, const int32_t* off_attn, const int32_t* off_mlp,
                          int32_t* rows_attn, int32_t* samp_attn, int32_t* rows_mlp, void* stream);
int laud_adavit_ln_rows(const float* x, int D, float eps, const float* w, const float* bias, const int32_t* row_idx,
                        const int32_t* row_cnt, int rows_max, void* y, void* stream);

/* Attention over the kept tokens (simulate_adavit.py:110-121: matmul + softmax + matmul on L_select tokens x kept heads).
 *   qkv fp16 [rows, ldq] compact rows of the attention list, HEAD-MAJOR columns: head h holds q | k | v (64 each) at
 *   column h*192; sample b owns rows [off[b], off[b+1]) (none: its attention sub-layer is skipped).
 *   o fp16 [rows, H*64]: o[i, h*64 : (h+1)*64] = softmax(q k^T / 8) v over the sample's rows for a kept head,
 *   zeros for a dropped head (so the projection of a sample with dropped heads stays exact).
 * head dimension 64; L = tokens per sample before selection (the bound of off[b+1] - off[b]), at most 208
 * (LAUD_E_UNSUPPORTED beyond: the score row of a query lives in registers). */
int laud_adavit_attention(const void* qkv, int ldq, const int32_t* off, const uint8_t* head_sel, int B, int H, int L,
                          void* o, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAUD_ADAVIT_H_ */
