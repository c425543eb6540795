// The product kernel of the mask-conditioned convolution: a persistent,
// warp-specialised gather-GEMM on the 5th-generation tensor cores (tcgen05).
//
//   D[m, n] = sum_{tap, k} A[m, (tap,k)] * W[n, (tap,k)]
//     m : up to 128 output pixels of ONE sample (or of a device-side row list)
//     n : output channels of one tile (<= 256)
//     k : that sample's ACTIVE input channels (compact), per filter tap
//
// 416 threads, one CTA per SM, grid = #SMs, static round-robin over work items
// (sample, m-tile, n-tile):
//   warps 0-3   epilogue : residual rows are prefetched into a shared staging tile
//                          with cp.async.bulk; the fp32 accumulator comes out of
//                          TMEM with tcgen05.ld; H1 pre-bias / folded BN / spatial
//                          gate / residual / ReLU in registers; the fp16 row is
//                          written back to the staging tile and leaves with one
//                          cp.async.bulk store per pixel row (compact channels);
//   warp  4     MMA      : one elected lane issues tcgen05.mma (M=128, runtime N,
//                          K=16) on 128B-swizzled shared-memory operands;
//                          accumulators double-buffered in TMEM (2 x 256 columns);
//   warps 5-12  producers: stage ONLY the active patches / channels with cp.async
//                          straight into the UMMA swizzle layout, signalling
//                          mbarriers (cp.async.mbarrier.arrive); stages are
//                          released by tcgen05.commit.  Three weight modes:
//        ROWS  : K-major B, row gather of the active OUTPUT channels (16-byte units);
//        KROWS : MN-major B from the transposed weights [tap][k][o]: row gather of the
//                active INPUT channels (16-byte units), all output channels of the
//                tile are computed and the epilogue compacts the active ones;
//        KUNITS: K-major B with an in-row gather of the active input channels
//                (4/8/16-byte units) - general fallback when no transposed copy
//                of the weights is supplied.
// Restates (does not port) the conv -> mask -> bn -> relu chains of
// imagenet_classification/models/laud_resnet.py:115-144 of the reference.
#include "laud_common.cuh"
#include "umma_ptx.cuh"

namespace laud {
namespace {

constexpr int BM = 128;                  // UMMA M: output pixels per tile
constexpr int BN_MAX = 256;              // UMMA N upper bound
constexpr int A_STAGE_BYTES = BM * 128;  // 128 rows x 64 fp16
constexpr int EPI_WARPS = 4, PROD_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int PROD_WARP0 = EPI_WARPS + 1;
constexpr int NUM_THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;   // 416
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int KIDX_MAX = 1024;
constexpr int MAX_STAGES = 8;
constexpr uint32_t TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 232448;               // 227 KB opt-in maximum per CTA

enum { MODE_ROWS = 1, MODE_KROWS = 2, MODE_KUNITS = 3 };

// Tap-validity rows for the H1-constant K-step: row `mask` (9-bit set of filter taps that fall
// inside the input) holds 1.0 at the valid taps, 0 elsewhere (16 fp16).  Filled once per process.
__device__ __half g_vtab[512 * 16];
__global__ void vtab_init_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 512 * 16) g_vtab[i] = __float2half(((i & 15) < 9 && ((i >> 4) >> (i & 15) & 1)) ? 1.f : 0.f);
}

struct Tables {
  int brow[BN_MAX];        // KUNITS: element offset of the weight row of compact column j (or -1)
  int kch[KIDX_MAX];       // real input channel of this sample's compact input channel e
  float scale[BN_MAX];
  float shift[BN_MAX];
  int ochan[BN_MAX];       // real output channel of tile column c (or -1: zero pad / inactive)
  int cpos[BN_MAX];        // position of tile column c in the stored row (staging / compact output)
  unsigned long long full[MAX_STAGES], empty[MAX_STAGES], tfull[2], tempty[2], rfull;
  uint32_t tmem_base;
};

struct Plan {              // host-computed launch geometry
  int n_mtiles, NT, total_items;
  int stages, stage_bytes, stg_pitch, mode;
  int staged;              // 1: rows leave through the shared staging tile + bulk copies
};

// ------------------------------------------------------------------ work items
struct Item {
  int b, mt, nt, ntiles, n0, n_valid, umma_n, Nc, Nfill, Kc, nk16, cpt, nchunks, has_bias;
};

__device__ __forceinline__ bool decode_item(const ConvArgs& a, const Plan& pl, int t, Item& it) {
  it.nt = t % pl.NT;
  const int r = t / pl.NT;
  it.mt = r % pl.n_mtiles;
  const int slot = r / pl.n_mtiles;
  int b = 0;
  if (a.row_idx) {
    if ((long long)it.mt * BM >= (long long)__ldg(a.row_cnt)) return false;
  } else {
    const int ns = a.sample_cnt ? __ldg(a.sample_cnt) : a.B;
    if (slot >= ns) return false;
    b = a.sample_idx ? __ldg(a.sample_idx + slot) : slot;
  }
  it.b = b;
  it.Nc = a.n_idx ? __ldg(a.n_cnt + b) * a.n_gran : a.C_out;
  it.Nfill = round_up(it.Nc, a.n_pad_align);
  // KROWS tiles span REAL output channels (the epilogue compacts); the others span the compact row
  const int span = (pl.mode == MODE_KROWS) ? a.C_out : it.Nfill;
  it.ntiles = (span + BN_MAX - 1) / BN_MAX;
  if (it.nt >= it.ntiles) return false;
  const int per = round_up((span + it.ntiles - 1) / it.ntiles, 16);
  it.n0 = it.nt * per;
  if (it.n0 >= span) return false;
  it.n_valid = min(per, span - it.n0);
  it.umma_n = round_up(it.n_valid, 16);
  it.Kc = a.k_idx ? __ldg(a.k_cnt + b) * a.k_gran : a.C_in;
  it.nk16 = (it.Kc + 15) >> 4;
  it.cpt = (it.nk16 + 3) >> 2;
  it.has_bias = (pl.mode == MODE_KROWS && a.bias_t != nullptr) ? 1 : 0;   // H1 constants as one extra K=16 step
  it.nchunks = it.cpt * a.ksize * a.ksize + it.has_bias;
  return true;
}

struct RowPos {
  int b, oy, ox, valid;
};
__device__ __forceinline__ RowPos row_pos(const ConvArgs& a, const Item& it, int r, int HWo) {
  RowPos p;
  const long long m = (long long)it.mt * BM + r;
  if (a.row_idx) {
    p.valid = m < (long long)__ldg(a.row_cnt);
    const int flat = p.valid ? __ldg(a.row_idx + m) : 0;
    p.b = flat / HWo;
    const int q = flat - p.b * HWo;
    p.oy = q / a.W_out;
    p.ox = q - p.oy * a.W_out;
  } else {
    p.valid = m < HWo;
    p.b = it.b;
    const int q = p.valid ? (int)m : 0;
    p.oy = q / a.W_out;
    p.ox = q - p.oy * a.W_out;
  }
  return p;
}

__device__ __forceinline__ int real_out_channel(const ConvArgs& a, int b, int jj) {
  return a.n_idx ? __ldg(a.n_idx + (size_t)b * a.n_ld + jj / a.n_gran) * a.n_gran + jj % a.n_gran : jj;
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_umma_kernel(const __grid_constant__ ConvArgs a,
                                                                   const __grid_constant__ Plan pl) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* stg = smem + (size_t)pl.stages * pl.stage_bytes;                       // staging tile [128][stg_pitch]
  Tables& T = *reinterpret_cast<Tables*>(stg + (size_t)(pl.staged ? BM * pl.stg_pitch : 0));
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HWo = a.H_out * a.W_out;
  const int taps = a.ksize * a.ksize;
  const int mode = pl.mode;
  if (a.row_idx) {            // density-dependent dispatch, decided on the device: not this kernel's regime -> nothing to do
    const int cnt = __ldg(a.row_cnt);
    if (cnt < a.row_lo || cnt >= a.row_hi) return;      // (uniform over the grid; before any barrier / TMEM allocation)
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < pl.stages; ++s) {
      mbar_init(&T.full[s], PROD_THREADS);
      mbar_init(&T.empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&T.tfull[i], 1);
      mbar_init(&T.tempty[i], EPI_THREADS);
    }
    mbar_init(&T.rfull, EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&T.tmem_base)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = T.tmem_base;

  if (warp >= PROD_WARP0) {
    // =========================================================== producers
    const int pt = threadIdx.x - PROD_WARP0 * 32;      // 0..255
    const int ac = pt & 7, ar0 = pt >> 3;              // 16-byte chunk, first row (rows ar0 + 32 i)
    const uint32_t off0 = sw128_off(ar0, ac);          // + 4096 i for row ar0 + 32 i
    // KUNITS geometry: bytes per unit, units per 128-byte row
    const int ub = (mode == MODE_KUNITS) ? min(16, a.k_gran * 2) : 16;
    const int upc = 128 / ub;
    const int ku = pt % upc, kr0 = pt / upc, krstep = PROD_THREADS / upc;      // krstep is a multiple of 8
    const uint32_t koff0 = sw128_off(kr0, (ku * ub) >> 4) + ((ku * ub) & 15);  // + 128 * krstep per step
    // KROWS geometry: warp pw stages k-rows pw + 8 i, lane = 16-byte chunk of the row
    const int pw = pt >> 5;
    const uint32_t roff0 = (uint32_t)((lane >> 3) * 8192 + pw * 128 + (((lane & 7) ^ pw) << 4));   // + 1024 i
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < pl.total_items; t += gridDim.x) {
      Item it;
      if (!decode_item(a, pl, t, it)) continue;
      named_bar_sync(1, PROD_THREADS);                 // previous item's table reads are done
      int browr[8];                                    // ROWS: weight-row offsets of the 8 rows this thread stages
      if (mode == MODE_ROWS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int jj = it.n0 + ar0 + 32 * i;
          browr[i] = (ar0 + 32 * i < it.umma_n && jj < it.Nc) ? real_out_channel(a, it.b, jj) * taps * a.C_in : -1;
        }
      } else if (mode == MODE_KUNITS) {
        for (int j = pt; j < it.umma_n; j += PROD_THREADS) {
          const int jj = it.n0 + j;
          T.brow[j] = jj < it.Nc ? real_out_channel(a, it.b, jj) * taps * a.C_in : -1;
        }
      }
      if (a.k_idx) {
        for (int e = pt; e < it.Kc; e += PROD_THREADS) {
          const int q = e / a.k_gran;
          T.kch[e] = __ldg(a.k_idx + (size_t)it.b * a.k_ld + q) * a.k_gran + (e - q * a.k_gran);
        }
      }
      // the 4 activation rows this thread stages: pixel base and top-left input coordinate
      int pbase[4], iy0[4], ix0[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const RowPos p = row_pos(a, it, ar0 + 32 * i, HWo);
        pbase[i] = p.b * a.H_in * a.W_in;
        iy0[i] = p.valid ? p.oy * a.stride - a.pad : -(1 << 20);
        ix0[i] = p.ox * a.stride - a.pad;
      }
      named_bar_sync(1, PROD_THREADS);
      const int ka_lim = a.k_idx ? it.nk16 * 16 : a.C_in;    // compact inputs are zero-padded to 16 by their producer

      for (int tap = 0; tap < taps; ++tap) {
        const int ty = tap / a.ksize, tx = tap - ty * a.ksize;
        const __half* aptr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int iy = iy0[i] + ty, ix = ix0[i] + tx;
          const bool ok = iy >= 0 && iy < a.H_in && ix >= 0 && ix < a.W_in;
          aptr[i] = ok ? a.x + (size_t)(pbase[i] + iy * a.W_in + ix) * a.ldx + ac * 8 : nullptr;
        }
        const int tapk = tap * a.C_in;
        for (int kq = 0; kq < it.cpt; ++kq) {
          const int k0 = kq * 64;
          const int n16 = min(4, it.nk16 - kq * 4);
          const int nch = 2 * n16;                            // 16-byte chunks the MMAs of this stage will read
          mbar_wait(&T.empty[stage], phase ^ 1);
          const uint32_t As = smem_base + stage * pl.stage_bytes;
          const uint32_t Bs = As + A_STAGE_BYTES;
          // ---- A: im2col gather of 128 pixel rows x 64 channels of this tap
          if (ac < nch) {
            const bool kok = k0 + ac * 8 < ka_lim;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const bool ok = kok && aptr[i] != nullptr;
              cp_async_16(As + off0 + 4096 * i, ok ? aptr[i] + k0 : a.x, ok ? 16u : 0u);
            }
          }
          // ---- B
          if (mode == MODE_ROWS) {
            if (ac < nch) {
              const int k = k0 + ac * 8;
              const bool kok = k < a.C_in;
              const __half* wk = a.w + tapk + k;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (ar0 + 32 * i < it.umma_n) {
                  const bool ok = kok && browr[i] >= 0;
                  cp_async_16(Bs + off0 + 4096 * i, ok ? wk + browr[i] : a.w, ok ? 16u : 0u);
                }
              }
            }
          } else if (mode == MODE_KROWS) {
            if (lane * 8 < it.umma_n) {
              const bool nok = it.n0 + lane * 8 < a.C_out;
              const __half* wn = a.wt + it.n0 + lane * 8;
              const int nrows = 16 * n16;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int kk = pw + 8 * i;
                if (kk < nrows) {
                  const int e = k0 + kk;
                  const bool ok = nok && e < it.Kc;
                  int rk = 0;
                  if (ok) rk = T.kch[e];
                  cp_async_16(Bs + roff0 + 1024 * i, ok ? wn + (size_t)(tapk + rk) * a.C_out : a.w, ok ? 16u : 0u);
                }
              }
            }
          } else if (ku * ub < nch * 16) {
            const int e = k0 + ku * (ub >> 1);                 // compact input channel of this unit
            const bool kok = e < it.Kc;
            int col = 0;
            if (kok) col = T.kch[e];
            const __half* wc = a.w + tapk + col;
            uint32_t dst = Bs + koff0;
            for (int j = kr0; j < it.umma_n; j += krstep, dst += 128 * krstep) {
              const int off = T.brow[j];
              const bool ok = kok && off >= 0;
              const __half* src = ok ? wc + off : a.w;
              if (ub == 4) cp_async_4(dst, src, ok ? 4u : 0u);
              else if (ub == 8) cp_async_8(dst, src, ok ? 8u : 0u);
              else cp_async_16(dst, src, ok ? 16u : 0u);
            }
          }
          cp_async_arrive(&T.full[stage]);
          if (++stage == pl.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (it.has_bias) {
        // one more K=16 step: A' = tap-validity indicator of each pixel, B' = this sample's H1 constants
        // T[b, tap, o] (zero rows for tap >= taps): adds sum_{valid taps} T[b,tap,o] to the accumulator
        mbar_wait(&T.empty[stage], phase ^ 1);
        const uint32_t As = smem_base + stage * pl.stage_bytes;
        const uint32_t Bs = As + A_STAGE_BYTES;
        if (ac < 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int vm = 0;
            for (int tap = 0; tap < taps; ++tap) {
              const int iy = iy0[i] + tap / a.ksize, ix = ix0[i] + tap % a.ksize;
              if (iy >= 0 && iy < a.H_in && ix >= 0 && ix < a.W_in) vm |= 1 << tap;
            }
            cp_async_16(As + off0 + 4096 * i, g_vtab + vm * 16 + ac * 8, 16u);
          }
        }
        if (lane * 8 < it.umma_n) {
          const bool nok = it.n0 + lane * 8 < a.C_out;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int kk = pw + 8 * i;
            const bool ok = nok && kk < taps;
            cp_async_16(Bs + roff0 + 1024 * i,
                        ok ? a.bias_t + (size_t)it.b * a.bias_ld + (size_t)kk * a.C_out + it.n0 + lane * 8 : a.w,
                        ok ? 16u : 0u);
          }
        }
        cp_async_arrive(&T.full[stage]);
        if (++stage == pl.stages) { stage = 0; phase ^= 1; }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == MMA_WARP) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < pl.total_items; t += gridDim.x) {
        Item it;
        if (!decode_item(a, pl, t, it)) continue;
        mbar_wait(&T.tempty[acc], aphase ^ 1);              // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN_MAX;
        const uint32_t idesc = umma_idesc_f16(it.umma_n, mode == MODE_KROWS);
        for (int ch = 0; ch < it.nchunks; ++ch) {
          const bool bias_step = it.has_bias && ch == it.nchunks - 1;
          const int n16 = bias_step ? 1 : min(4, it.nk16 - (ch % it.cpt) * 4);
          mbar_wait(&T.full[stage], phase);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t As = smem_base + stage * pl.stage_bytes;
          const uint32_t Bs = As + A_STAGE_BYTES;
          const uint64_t ad = umma_desc(As, 16, 1024);
          if (mode == MODE_KROWS) {
            // B rows are k: one MMA consumes two 8-k groups (2 x 1024 B); 64-n blocks are 8192 B apart
            const uint64_t bd = umma_desc(Bs, 8192, 1024);
            for (int k = 0; k < n16; ++k) umma_f16(d_tmem, ad + 2 * k, bd + 128 * k, idesc, (ch | k) ? 1u : 0u);
          } else {
            const uint64_t bd = umma_desc(Bs, 16, 1024);
            for (int k = 0; k < n16; ++k)                      // +32 B per K=16 step inside the swizzle atom
              umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (ch | k) ? 1u : 0u);
          }
          umma_commit(&T.empty[stage]);                         // frees the stage when these MMAs retire
          if (++stage == pl.stages) { stage = 0; phase ^= 1; }
        }
        if (it.nchunks > 0) umma_commit(&T.tfull[acc]);
        else mbar_arrive(&T.tfull[acc]);                        // no active input channel: accumulator is all zero
        if (++acc == 2) { acc = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =========================================================== epilogue
    const int et = threadIdx.x;                                 // 0..127 == accumulator row (TMEM lane)
    unsigned char* srow = stg + (size_t)et * pl.stg_pitch;      // this thread's row of the staging tile
    const uint32_t srow_u32 = smem_u32(srow);
    const bool compact = (mode == MODE_KROWS) && a.n_idx;       // tile columns are real channels; store the active ones
    int acc = 0;
    uint32_t aphase = 0, rphase = 0;
    for (int t = blockIdx.x; t < pl.total_items; t += gridDim.x) {
      Item it;
      if (!decode_item(a, pl, t, it)) continue;
      named_bar_sync(2, EPI_THREADS);                           // previous item's table reads are done
      for (int c = et; c < it.umma_n; c += EPI_THREADS) {
        const int jj = it.n0 + c;
        int o = -1, pos = c;
        if (compact) {
          // real channel jj: active iff its group is in the sample's ascending list; its rank is the compact position
          if (c < it.n_valid) {
            const int grp = jj / a.n_gran, na = it.Nc / a.n_gran;
            const int* lst = a.n_idx + (size_t)it.b * a.n_ld;
            int lo = 0, hi = na;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (__ldg(lst + mid) < grp) lo = mid + 1; else hi = mid;
            }
            if (lo < na && __ldg(lst + lo) == grp) { o = jj; pos = lo * a.n_gran + jj % a.n_gran; }
          }
        } else if (jj < it.Nc) {
          o = real_out_channel(a, it.b, jj);
        }
        float sc = o >= 0 ? 1.f : 0.f, sh = 0.f;      // inactive / pad columns come out as exact zeros
        if (o >= 0 && a.scale) { sc = __ldg(a.scale + o); sh = __ldg(a.shift + o); }
        if (o >= 0 && a.n_mask && a.n_mask[(size_t)it.b * (a.C_out / a.n_mask_gran) + o / a.n_mask_gran] == 0) sc = 0.f;
        T.ochan[c] = o; T.cpos[c] = o >= 0 ? pos : -1; T.scale[c] = sc; T.shift[c] = sh;
      }
      const RowPos p = row_pos(a, it, et, HWo);
      const size_t pix = (size_t)p.b * HWo + (size_t)p.oy * a.W_out + p.ox;
      int cls = 0;
      if (a.pre_bias && a.pre_bias_classes > 1)
        cls = border_class(p.oy, a.stride, a.pad, a.H_in) * 4 + border_class(p.ox, a.stride, a.pad, a.W_in);
      const float* pb = a.pre_bias ? a.pre_bias + ((size_t)p.b * a.pre_bias_classes + cls) * a.pre_bias_ld + it.n0 : nullptr;
      const int cpg = a.out_mask ? a.C_out / a.mask_groups : 1;
      // stored row of this tile: [row_c0, row_c0 + row_n) of y's channel axis
      const int row_c0 = compact ? 0 : it.n0;
      const int row_n = compact ? it.Nfill : it.n_valid;
      const bool staged = pl.staged && !(compact && it.ntiles > 1);
      if (pl.staged) bulk_wait_read();                          // the previous store of this row has left shared memory
      if (staged && a.residual) {
        if (p.valid) {
          mbar_arrive_expect_tx(&T.rfull, (uint32_t)row_n * 2);
          bulk_load(srow_u32, a.residual + pix * a.ldr + row_c0, (uint32_t)row_n * 2, &T.rfull);
        } else {
          mbar_arrive(&T.rfull);
        }
      }
      named_bar_sync(2, EPI_THREADS);
      mbar_wait(&T.tfull[acc], aphase);
      tc_fence_after();
      if (staged && a.residual) { mbar_wait(&T.rfull, rphase); rphase ^= 1; }
      const uint32_t taddr = tmem_base + acc * BN_MAX + ((uint32_t)(warp * 32) << 16);
      // generic path: per-class pre-bias table, per-group spatial gates, RELU_WHERE_GATE0 (fallback layouts)
      const bool generic = pb != nullptr || (a.out_mask && a.mask_groups != 1) || a.relu_mode == LAUD_RELU_WHERE_GATE0;
      const bool relu_all = a.relu_mode == LAUD_RELU_ALL;
      bool row_on = true;                                        // spatial gate of this pixel (one mask group)
      if (!generic && a.out_mask && p.valid) row_on = a.out_mask[(size_t)p.b * HWo + (size_t)p.oy * a.W_out + p.ox] != 0;
      for (int c0 = 0; c0 < it.n_valid; c0 += 16) {
        float v[16];
        if (it.nchunks > 0) {
          tmem_ld16(taddr + c0, v);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0.f;
        }
        if (!p.valid) continue;
        __align__(16) __half2 out2[8];
        __half* out = reinterpret_cast<__half*>(out2);
        __align__(16) __half2 rs2[8];
        if (a.residual) {
          if (staged) {
            *reinterpret_cast<uint4*>(rs2) = *reinterpret_cast<const uint4*>(srow + c0 * 2);
            *reinterpret_cast<uint4*>(rs2 + 4) = *reinterpret_cast<const uint4*>(srow + c0 * 2 + 16);
          } else {
            const uint4* rp = reinterpret_cast<const uint4*>(a.residual + pix * a.ldr + it.n0 + c0);
            *reinterpret_cast<uint4*>(rs2) = __ldg(rp);
            if (c0 + 8 < it.n_valid) *reinterpret_cast<uint4*>(rs2 + 4) = __ldg(rp + 1);
          }
        }
        if (!generic) {
          const float4* scp = reinterpret_cast<const float4*>(T.scale + c0);
          const float4* shp = reinterpret_cast<const float4*>(T.shift + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 s4 = scp[q], h4 = shp[q];
            v[4 * q] = fmaf(v[4 * q], s4.x, h4.x);
            v[4 * q + 1] = fmaf(v[4 * q + 1], s4.y, h4.y);
            v[4 * q + 2] = fmaf(v[4 * q + 2], s4.z, h4.z);
            v[4 * q + 3] = fmaf(v[4 * q + 3], s4.w, h4.w);
          }
          if (!row_on) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0.f;
          }
          if (a.residual) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float2 r = __half22float2(rs2[q]);
              v[2 * q] += r.x;
              v[2 * q + 1] += r.y;
            }
          }
          if (relu_all) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) out2[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        } else {
          if (pb) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 f = __ldg(reinterpret_cast<const float4*>(pb + c0) + q);
              v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w;
            }
          }
          const __half* rs = reinterpret_cast<const __half*>(rs2);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int o = T.ochan[c0 + e];
            float val = v[e] * T.scale[c0 + e] + T.shift[c0 + e];
            int gate = 1;
            if (a.out_mask && o >= 0) {
              gate = a.out_mask[((size_t)p.b * a.mask_groups + o / cpg) * HWo + (size_t)p.oy * a.W_out + p.ox];
              if (a.relu_mode != LAUD_RELU_WHERE_GATE0) val = gate ? val : 0.f;
            }
            if (a.residual) val += __half2float(rs[e]);
            if (a.relu_mode == LAUD_RELU_ALL || (a.relu_mode == LAUD_RELU_WHERE_GATE0 && !gate)) val = fmaxf(val, 0.f);
            out[e] = __float2half(o >= 0 ? val : 0.f);
          }
        }
        if (compact) {
          // scatter the active columns to their compact positions (2-byte stores)
          const int4* cpp = reinterpret_cast<const int4*>(T.cpos + c0);
          __half* dst = staged ? reinterpret_cast<__half*>(srow) : a.y + pix * a.ldy;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int4 cp = cpp[q];
            if (cp.x >= 0) dst[cp.x] = out[4 * q];
            if (cp.y >= 0) dst[cp.y] = out[4 * q + 1];
            if (cp.z >= 0) dst[cp.z] = out[4 * q + 2];
            if (cp.w >= 0) dst[cp.w] = out[4 * q + 3];
          }
        } else if (staged) {
          *reinterpret_cast<uint4*>(srow + c0 * 2) = *reinterpret_cast<const uint4*>(out2);
          *reinterpret_cast<uint4*>(srow + c0 * 2 + 16) = *reinterpret_cast<const uint4*>(out2 + 4);
        } else {
          __half* yp = a.y + pix * a.ldy + it.n0 + c0;
          *reinterpret_cast<uint4*>(yp) = *reinterpret_cast<const uint4*>(out2);
          if (c0 + 8 < it.n_valid) *reinterpret_cast<uint4*>(yp + 8) = *reinterpret_cast<const uint4*>(out2 + 4);
        }
      }
      tc_fence_before();
      mbar_arrive(&T.tempty[acc]);                              // accumulator drained: the MMA warp may reuse it
      if (++acc == 2) { acc = 0; aphase ^= 1; }
      if (p.valid && compact) {
        // zero pad [Nc, Nfill) of the compact row (once per row: by the last n-tile)
        if (staged) {
          for (int j = it.Nc; j < it.Nfill; ++j) *reinterpret_cast<__half*>(srow + j * 2) = __float2half(0.f);
        } else if (it.nt == it.ntiles - 1) {
          for (int j = it.Nc; j < it.Nfill; ++j) a.y[pix * a.ldy + j] = __float2half(0.f);
        }
      }
      if (staged) {
        fence_proxy_async();                                    // generic-proxy row writes -> visible to the bulk copy
        if (p.valid && row_n > 0) bulk_store(a.y + pix * a.ldy + row_c0, srow_u32, (uint32_t)row_n * 2);
        bulk_commit();
      }
    }
    bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Layouts the tcgen05 kernel does not take (they go to the legacy HMMA kernel):
// channel granularity 1 (2-byte gather units), unaligned pitches, odd widths.
bool conv_umma_supported(const ConvArgs& a) {
  if (a.C_in % 8 || a.ldx % 8 || a.ldy % 8 || a.C_out % 8) return false;
  if (!aligned16(a.x) || !aligned16(a.w) || !aligned16(a.y)) return false;
  if (a.residual && (a.ldr % 8 || !aligned16(a.residual))) return false;
  if (a.pre_bias && (a.pre_bias_ld % 4 || !aligned16(a.pre_bias))) return false;
  if (a.k_idx) {
    if (a.k_gran != 2 && a.k_gran != 4 && a.k_gran % 8) return false;   // 4-, 8- or 16-byte gather units
    if (a.C_in > KIDX_MAX) return false;
    if (a.ldx < round_up(a.C_in, 16)) return false;
  }
  if (a.n_idx && a.n_pad_align < 8) return false;
  if ((size_t)a.C_out * a.ksize * a.ksize * a.C_in >= (1ull << 31)) return false;
  return true;
}

int conv_forward_umma(const ConvArgs& a, cudaStream_t s) {
  if (!conv_umma_supported(a)) return conv_forward_hmma(a, s);
  static int num_sms_dev[MAX_DEVICES] = {0};
  const int dev = current_device();
  if (num_sms_dev[dev] == 0) {
    int n = 0;
    LAUD_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    LAUD_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    vtab_init_kernel<<<32, 256, 0, s>>>();      // tap-validity table of the H1-constant K-step (once per device)
    if (int e = check_launch("vtab_init_kernel")) return e;
    if (int e = finish_first_call_init(s, "conv_forward_umma")) return e;
    num_sms_dev[dev] = n;
  }
  const int num_sms = num_sms_dev[dev];
  Plan pl;
  pl.mode = (a.k_idx && a.wt && aligned16(a.wt)) ? MODE_KROWS : (a.k_idx ? MODE_KUNITS : MODE_ROWS);
  const long long HWo = (long long)a.H_out * a.W_out;
  const long long rows = a.row_idx ? (long long)a.B * HWo : HWo;
  pl.n_mtiles = (int)((rows + BM - 1) / BM);
  const int nfill_max = round_up(a.C_out, a.n_pad_align);
  const int span = pl.mode == MODE_KROWS ? a.C_out : nfill_max;
  pl.NT = (span + BN_MAX - 1) / BN_MAX;
  const long long total = (long long)(a.row_idx ? 1 : a.B) * pl.n_mtiles * pl.NT;
  if (total >= (1ll << 31)) {
    set_error("conv_forward_umma: too many tiles");
    return LAUD_E_BADARG;
  }
  pl.total_items = (int)total;
  // widest tile of this launch -> B stage size, staging row pitch, pipeline depth
  const int bn_cap = min(BN_MAX, round_up(span, 16));
  const int b_bytes = pl.mode == MODE_KROWS ? ((bn_cap + 63) / 64) * 8192 : bn_cap * 128;
  pl.stage_bytes = A_STAGE_BYTES + round_up(b_bytes, 1024);
  const int row_max = (pl.mode == MODE_KROWS && a.n_idx) ? max(round_up(nfill_max, 16), bn_cap) : bn_cap;   // stored-row width
  pl.stg_pitch = row_max * 2 + 16;
  pl.staged = 1;
  int avail = SMEM_LIMIT - 1024 - (int)sizeof(Tables) - BM * pl.stg_pitch;
  if (avail < 2 * pl.stage_bytes) {        // staging tile does not fit beside a 2-stage pipeline: direct stores
    pl.staged = 0;
    avail = SMEM_LIMIT - 1024 - (int)sizeof(Tables);
  }
  pl.stages = avail / pl.stage_bytes;
  if (pl.stages > MAX_STAGES) pl.stages = MAX_STAGES;
  const size_t smem = 1024 + (size_t)pl.stages * pl.stage_bytes + (pl.staged ? (size_t)BM * pl.stg_pitch : 0) + sizeof(Tables);
  const int grid = (int)(total < num_sms ? total : num_sms);
  g_conv_paths[0].fetch_add(1, std::memory_order_relaxed);
  {
    ConvProfScope prof(s);
    conv_umma_kernel<<<grid, NUM_THREADS, smem, s>>>(a, pl);
  }
  return check_launch("conv_umma_kernel");
}

}  // namespace laud
