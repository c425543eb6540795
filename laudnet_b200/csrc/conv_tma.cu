// The product kernel of the mask-conditioned convolution: a persistent, warp-specialised
// implicit GEMM on tcgen05 tensor cores whose dense operands move with TMA tensor-tile copies.
//
//   D[m, n] = sum_{tap, k} A[m, (tap,k)] * W[n, (tap,k)]
//     m : output pixels: MT (1|2) tiles of up to 128 pixels share every weight stage.  Tiles belong
//         to ONE sample when anything is per sample (channel lists / gates, sample lists, 3x3);
//         a 1x1 layer with nothing per sample is one FLAT GEMM over all B*H*W pixels.
//     n : output channels of one tile (BN = 64 | 128 | 256)
//     k : input channels per filter tap (the sample's ACTIVE ones, compact, in the gathered modes)
//
// 512 threads, one CTA per SM, one contiguous range of work items (sample, m-group, n-group) per CTA:
//   warps 0-7   epilogue : tcgen05.ld the fp32 accumulators out of TMEM; folded BN / gates /
//                          residual / ReLU in registers.  Column tables (scale, shift) are loaded
//                          once per CTA when they do not depend on the sample.
//                          SLAB mode: each half (4 warps) owns a ring of [128 px][64 ch] swizzled
//                          slabs; residual slabs ARRIVE by TMA (kept ring-1 tasks ahead), the fp16
//                          result overwrites them in place and LEAVES by TMA.
//                          DIRECT mode (long reductions, no residual): registers -> global.
//                          ROWS mode (K-row gather + compaction of the active output channels):
//                          word-swizzled row staging, coalesced cooperative flush.
//   warp  8     TMA      : one lane issues the activation tiles - 3-d boxes for 1x1 layers, 4-d boxes
//                          for 3x3 / stride-2 layers (im2col and subsampling by the TMA unit, taps
//                          that fall outside the image zero-filled) - and the weight tiles when they
//                          are not gathered.  HALO mode (3x3, stride 1, shared weights): the
//                          activations of an m-group are staged ONCE per 64-channel chunk with their
//                          zero halo; the nine taps are that tile read at nine row offsets.
//   warp  9     MMA      : the whole warp runs the issue loop with identical values and an elected lane
//                          issues tcgen05.mma (M=128, N<=256, K=16) - issued from a lane-0 region every
//                          UTCHMMA is wrapped in an ELECT / R2UR / BRA.U.ANY loop that costs as much as the
//                          MMA; accumulators in TMEM, multi-buffered so the epilogue of one item overlaps
//                          the MMAs of the next.
//   warps 10-15 gather   : per-sample weight gathers with cp.async into the UMMA swizzle layout
//                          (ROWS: active OUTPUT channels of K-major weights; KROWS: active INPUT
//                          channels of the transposed weights) + the H1-constant K=16 step.
//                          With shared weights two of their lanes are the slab DMA threads: they
//                          issue the TMA stores / residual loads of the two epilogue halves; four of the
//                          warps are the POOLING warps of the fused global average pool (gap_partial):
//                          they add up the columns of every finished output slab per sample, so the next
//                          block's channel masker never reads the activations.
// The generic kernel (SPEC 0) is ~160 KB of code; the shared-weight layers run on instantiations with their
// modes fixed at compile time (Mode<SPEC>): instruction-cache misses cost every role 0.5-1.4k cycles per item.
// Layouts it does not take (row lists, per-class pre-bias, KUNITS gathers) run on the cp.async-staged
// kernel (conv_umma.cu).  Restates (does not port) laud_resnet.py:115-144 of the reference.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "laud_common.cuh"
#include "umma_ptx.cuh"

namespace laud {
std::atomic<int> g_conv_pdl{0};
namespace {

constexpr int BM = 128;
constexpr int A_TILE_BYTES = BM * 128;           // 128 rows x 64 fp16
constexpr int EPI_WARPS = 8, TMA_WARP = 8, MMA_WARP = 9, GATHER_WARP0 = 10, GATHER_WARPS = 6;
constexpr int MMA_WARP2 = 15;                     // second MMA issuer (Plan::dual): another scheduler than MMA_WARP's
constexpr int NUM_THREADS = (GATHER_WARP0 + GATHER_WARPS) * 32;   // 512
constexpr int EPI_THREADS = EPI_WARPS * 32, HALF_THREADS = EPI_THREADS / 2;
constexpr int GATHER_THREADS = GATHER_WARPS * 32;
constexpr int KIDX_MAX = 1024;
constexpr int MAX_STAGES = 6;
constexpr int SLAB_BYTES = BM * 128;             // [128 px][64 ch] fp16
constexpr int MAX_RING = 3;
constexpr int BN_MAX = 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 232448;
constexpr int XW_MAX = 512;                      // OUT_EXPAND: channel pairs of a dense row (C_out <= 1024)
constexpr int G4_BN = 192;                       // gather4 path: compact columns per n-tile (a 24 KB stage; typical active
                                                 // counts of a 256-channel layer at density 0.6 fit one tile)

enum { BMODE_TMA = 0, BMODE_ROWS = 1, BMODE_KROWS = 2, BMODE_G4 = 3 };   // G4: active weight rows by TMA tile::gather4
enum { OUT_SLAB = 0, OUT_ROWS = 1, OUT_DIRECT = 2, OUT_EXPAND = 3 };     // EXPAND: compact columns -> dense rows + BN constants

// Optional in-kernel lap timers (-DLAUD_KPROF, scripts/kprof.py): cycles each warp role spends per phase.
#ifdef LAUD_KPROF
constexpr int KP_ROLES = 5, KP_SITES = 8;
__device__ long long g_kprof[160 * KP_ROLES * KP_SITES];
#define KP_DECL long long kp_t = clock64(), kp_acc[KP_SITES] = {0, 0, 0, 0, 0, 0, 0, 0}
#define KP_LAP(i) do { const long long kp_n = clock64(); kp_acc[i] += kp_n - kp_t; kp_t = kp_n; } while (0)
#define KP_FLUSH(role) do { if (blockIdx.x < 160) for (int kp_i = 0; kp_i < KP_SITES; ++kp_i) \
    g_kprof[(blockIdx.x * KP_ROLES + (role)) * KP_SITES + kp_i] = kp_acc[kp_i]; } while (0)
#else
#define KP_DECL
#define KP_LAP(i)
#define KP_FLUSH(role)
#endif

__device__ __half g_vtab4[512 * 16];             // tap-validity rows of the H1-constant K-step (see conv_umma.cu)
__global__ void vtab4_init_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 512 * 16) g_vtab4[i] = __float2half(((i & 15) < 9 && ((i >> 4) >> (i & 15) & 1)) ? 1.f : 0.f);
}

constexpr int CNT_CACHE = 512;                   // per-sample active counts cached in shared memory when B fits

struct alignas(16) Tables {
  int kch[KIDX_MAX];       // real input channel of this sample's compact input channel e
  float scale[2][BN_MAX];  // epilogue column tables, double-buffered across sub-items
  float shift[2][BN_MAX];
  int cpos[2][BN_MAX];     // OUT_ROWS: compact position of tile column c, or -1
  int kcnt[CNT_CACHE], ncnt[CNT_CACHE];
  unsigned short vmask[2][BM];   // tap-validity bit sets of the pixels of the current m-group (H1-constant step)
  unsigned long long full[MAX_STAGES], empty[MAX_STAGES], tfull[4], tempty[4], rfull[2][MAX_RING];
  unsigned long long afull[2], aempty[2];   // halo mode: ring of activation tiles (one per 64-channel chunk)
  unsigned long long sready[2][MAX_RING], sfree[2][MAX_RING];   // slab hand-off between the epilogue halves and their DMA threads
  unsigned long long gdone[2][MAX_RING];    // fused GAP: the pooling warps have read the slab
  uint32_t tmem_base;
  // OUT_EXPAND: per sample, for every 32-bit word (channel pair) of a dense output row: the word of the compact
  // staging row that holds it, or -1 = gated -> the BN constant pair cw
  short wsrc[XW_MAX];
  uint32_t cw[XW_MAX];
};

struct Plan {              // host-computed launch geometry
  int n_mtiles, n_mgroups, MT, NT, NTI, BN, total_items;
  int stages, stage_bytes, b_off;
  int bmode, omode;
  int R;                   // 3x3: output rows per m-tile (0 for 1x1)
  int rows_per_tile;       // pixels of a full m-tile: 128 (1x1) or R*W_out
  int nbuf, acc_cols;      // accumulator buffers; TMEM columns of one accumulator
  int ring;                // slabs per half (OUT_SLAB)
  int stg_pitch, stg_rows; // OUT_ROWS / OUT_EXPAND: bytes per staging row, rows
  int stg_bytes;           // bytes of the slab rings / row staging in front of the tables
  int full_count;          // arrivals that complete a stage
  int a_tx, b_tx, r_tx;    // bytes one A-tile / B-tile / residual-slab TMA copy delivers
  int cnt_cached;          // 1: Tables::kcnt / ncnt hold k_cnt / n_cnt of every sample
  // HALO mode (3x3, stride 1, shared weights): the activations of an m-group are staged ONCE per 64-channel chunk,
  // with their zero-padded 1-pixel halo, as rows of a (W+2)-wide padded image; the nine taps are the same tile read
  // at nine row offsets (UMMA descriptors with a base offset), so only the weights stream per tap.
  int halo, Wp, a_slot_bytes, a_region_bytes, halo_bo;
  // layers whose output columns do not depend on the sample (no gather, no channel gate): folded-BN scale / shift of ALL
  // output channels are loaded once per CTA (stab_cols floats each, zero padded) - no per-item column tables, no barrier
  int static_cols, stab_cols;
  int dma;                 // 1: slab stores / residual prefetches are issued by two otherwise idle threads (shared-weight layers)
  int simple;              // 1: nothing per sample (no lists, gathers): Sub fields below are launch constants
  int c_nfill, c_nk16, c_cpt, c_nchunks, NG;
  int st256;               // 1: OUT_DIRECT may use 32-byte global stores (y 32-byte aligned, ldy and C_out multiples of 16)
  int gap;                 // 1: fused global-average-pool partial sums of the output (flat 1x1 layers, OUT_SLAB + dma)
  int flat;                // 1: the layer runs as ONE GEMM over all B*H*W pixels (a.B == 1, a.gap_hw = pixels per sample)
  int nm;                  // 1: masked-dense channel gate (n_mask) looked up per accumulator ROW in the epilogue - the column
                           //    tables stay static, so a gated 1x1 layer can still be flat;  nm_fast: granularity 2, 16-byte rows
  int nm_fast;
  int dual;                // 1: TWO MMA-issuing warps (MT == 2, direct / expanding epilogue): warp 9 issues the MMAs of m-tile 0,
                           //    warp 15 those of m-tile 1.  An issued tcgen05.mma costs ~24 issue slots (seven R2UR.BROADCAST
                           //    moves of its operands into uniform registers) on a scheduler shared with two epilogue warps:
                           //    ~110 cycles per MMA from one warp whatever N (measured, profiles/r02*_kprof*), i.e. every
                           //    layer with N < 256 - and channel skipping - was issue-bound, not tensor-bound
  int g4_cp;               // BMODE_G4: 1 = the active weight rows are staged by the six gather warps with 16-byte cp.async
                           //   (192 threads, ~7 copies each per stage); 0 = by TMA tile::gather4 from the producer warp
                           //   (four rows per instruction - measured issue-bound: ~100 cycles per copy, profiles/r02c_*)
  int tsplit;              // 1: (flat, one n-group, MT == 2) CTAs own contiguous ranges of m-TILES, processed two at a time:
                           //    the critical path is ceil(tiles / CTAs) tiles instead of 2 x ceil(tile pairs / CTAs)
  int dbg;                 // LAUD_DBG timing experiments (wrong results): 2 no activation loads, 4 no MMAs, 8 no epilogue work, 16 half-N MMAs
};

struct Sub {               // one (sample, m-group, n-tile) unit of work
  int b, mg, nt, n0, n_valid, umma_n, Nc, Nfill, Kc, nk16, cpt, nchunks, has_bias, mt0, mt_cnt;
};

// Kernel specialisations.  The generic kernel (SPEC 0) carries every mode; its code is far larger than the instruction
// caches, and each warp role pays for that at every item.  The shared-weight layers with nothing per sample - all
// convolutions of the masked-dense network - run on instantiations with their modes fixed at compile time:
//   1: OUT_SLAB through the DMA threads   2: OUT_DIRECT (conv1)   3: OUT_DIRECT, halo tiles (3x3)
//   4: as 1 with residual + ReLU and no pixel gate (conv3)   5: as 1 without residual / ReLU / gate (downsample)
template <int SPEC>
struct Mode {
//   6: channel skipping with a dense result (n_expand): halo tiles, the sample's ACTIVE weight rows by TMA gather4,
//      compact accumulator columns expanded to dense rows in the epilogue
  static constexpr bool SLAB_FIXED = SPEC == 4 || SPEC == 5;
  static constexpr bool G4 = SPEC == 6;
  static __device__ __forceinline__ int bmode(const Plan& pl) { return SPEC == 0 ? pl.bmode : (G4 ? (int)BMODE_G4 : (int)BMODE_TMA); }
  static __device__ __forceinline__ int omode(const Plan& pl) {
    return SPEC == 0 ? pl.omode : (G4 ? (int)OUT_EXPAND : ((SPEC == 1 || SLAB_FIXED) ? (int)OUT_SLAB : (int)OUT_DIRECT));
  }
  static __device__ __forceinline__ bool halo(const Plan& pl) { return SPEC == 0 ? pl.halo != 0 : (SPEC == 3 || G4); }
  static __device__ __forceinline__ bool dma(const Plan& pl) { return SPEC == 0 ? pl.dma != 0 : (SPEC == 1 || SLAB_FIXED); }
  static __device__ __forceinline__ bool simple(const Plan& pl) { return SPEC == 0 ? pl.simple != 0 : !G4; }
  static __device__ __forceinline__ bool has_res(const ConvArgs& a) { return SLAB_FIXED ? SPEC == 4 : a.residual != nullptr; }
  static __device__ __forceinline__ int relu_mode(const ConvArgs& a) {
    return SLAB_FIXED ? (SPEC == 4 ? (int)LAUD_RELU_ALL : (int)LAUD_RELU_NONE) : a.relu_mode;
  }
  static __device__ __forceinline__ const uint8_t* out_mask(const ConvArgs& a) { return SLAB_FIXED ? nullptr : a.out_mask; }
};

// Every role walks the same CONTIGUOUS range of items (sample slot, m-group, n-group) of its CTA, and the
// n-tiles inside an item, in the same order; consecutive sub-items mostly share the sample.
struct Walker {
  int t, t_end, slot, mg, ng, nti, started;
};
__device__ __forceinline__ void walker_init(const ConvArgs& a, const Plan& pl, Walker& w) {
  if (pl.tsplit) {                           // t counts m-tiles; an item is up to MT consecutive tiles of the CTA's range
    const int base = pl.n_mtiles / (int)gridDim.x, rem = pl.n_mtiles % (int)gridDim.x;
    const int c = (int)blockIdx.x;
    w.t = c * base + min(c, rem);
    w.t_end = w.t + base + (c < rem ? 1 : 0);
    w.ng = 0; w.mg = 0; w.slot = 0; w.nti = 0; w.started = 0;
    return;
  }
  // with a device-side sample list only the ACTIVE samples' items are partitioned, so every CTA gets its share
  const int total = a.sample_cnt ? min(__ldg(a.sample_cnt), a.B) * (pl.total_items / a.B) : pl.total_items;
  const int base = total / (int)gridDim.x, rem = total % (int)gridDim.x;
  const int c = (int)blockIdx.x;
  w.t = c * base + min(c, rem);
  w.t_end = w.t + base + (c < rem ? 1 : 0);
  const int NG = pl.NG;
  w.ng = w.t % NG;
  const int r = w.t / NG;
  w.mg = r % pl.n_mgroups;
  w.slot = r / pl.n_mgroups;
  w.nti = 0;
  w.started = 0;
}
template <class M>
__device__ __forceinline__ bool decode_sub(const ConvArgs& a, const Plan& pl, const Tables& T, const Walker& w, Sub& s) {
  if (M::simple(pl)) {                     // division-free path of the shared-weight layers
    s.b = a.sample_idx ? __ldg(a.sample_idx + w.slot) : w.slot;      // (walker_init hands out the ACTIVE samples' items only)
    s.mg = w.mg;
    s.nt = w.ng * pl.NTI + w.nti;
    s.Nc = a.C_out;
    s.Nfill = pl.c_nfill;
    s.n0 = s.nt * pl.BN;
    if (s.n0 >= s.Nfill) return false;
    s.n_valid = min(pl.BN, s.Nfill - s.n0);
    s.umma_n = (s.n_valid + 15) & ~15;
    s.Kc = a.C_in;
    s.nk16 = pl.c_nk16;
    s.cpt = pl.c_cpt;
    s.has_bias = 0;
    s.nchunks = pl.c_nchunks;
    s.mt0 = w.mg * pl.MT;
    s.mt_cnt = min(pl.MT, pl.n_mtiles - s.mt0);
    if (pl.tsplit) {
      s.mt0 = w.t;
      s.mt_cnt = min(pl.MT, w.t_end - w.t);
    }
    return true;
  }
  const int ns = a.sample_cnt ? __ldg(a.sample_cnt) : a.B;
  if (w.slot >= ns) return false;
  s.b = a.sample_idx ? __ldg(a.sample_idx + w.slot) : w.slot;
  s.mg = w.mg;
  s.nt = w.ng * pl.NTI + w.nti;
  s.Nc = a.n_idx ? (pl.cnt_cached ? T.ncnt[s.b] : __ldg(a.n_cnt + s.b)) * a.n_gran : a.C_out;
  s.Nfill = round_up(s.Nc, a.n_pad_align);
  // KROWS tiles span REAL output channels (the epilogue compacts); the others span the stored row
  // (gather4 path: a sample without any active channel still gets one 16-column tile - its rows are all constants)
  const int span = (M::bmode(pl) == BMODE_KROWS) ? a.C_out : (M::bmode(pl) == BMODE_G4 ? max(s.Nfill, 16) : s.Nfill);
  s.n0 = s.nt * pl.BN;
  if (s.n0 >= span) return false;
  s.n_valid = min(pl.BN, span - s.n0);
  s.umma_n = round_up(s.n_valid, 16);
  s.Kc = a.k_idx ? (pl.cnt_cached ? T.kcnt[s.b] : __ldg(a.k_cnt + s.b)) * a.k_gran : a.C_in;
  s.nk16 = (s.Kc + 15) >> 4;
  s.cpt = (s.nk16 + 3) >> 2;
  s.has_bias = a.bias_t != nullptr ? 1 : 0;
  s.nchunks = s.cpt * a.ksize * a.ksize + s.has_bias;
  s.mt0 = w.mg * pl.MT;
  s.mt_cnt = min(pl.MT, pl.n_mtiles - s.mt0);
  return true;
}
// advance to the next valid sub-item of this CTA; false when the range is exhausted
template <class M>
__device__ __forceinline__ bool walker_next(const ConvArgs& a, const Plan& pl, const Tables& T, Walker& w, Sub& s) {
  while (true) {
    if (!w.started) {
      w.started = 1;
    } else if (pl.tsplit) {
      w.t += pl.MT;
    } else if (++w.nti >= pl.NTI) {
      w.nti = 0;
      ++w.t;
      if (++w.ng >= pl.NG) {
        w.ng = 0;
        if (++w.mg >= pl.n_mgroups) { w.mg = 0; ++w.slot; }
      }
    }
    if (w.t >= w.t_end) return false;
    if (decode_sub<M>(a, pl, T, w, s)) return true;
  }
}

// first output pixel (within the sample) and number of valid pixels of m-tile mt
__device__ __forceinline__ void tile_rows(const ConvArgs& a, const Plan& pl, int mt, int& m0, int& rows) {
  if (pl.R) {
    const int oy0 = mt * pl.R;
    m0 = oy0 * a.W_out;
    rows = min(pl.R, a.H_out - oy0) * a.W_out;
  } else {
    m0 = mt * BM;
    rows = min(BM, a.H_out * a.W_out - m0);
  }
}

// slab tasks of one epilogue half, in execution order (used by the residual prefetcher)
struct Cursor {
  Walker w;
  Sub s;
  int mt, sl, have;
};
template <class M>
__device__ __forceinline__ bool cursor_next(const ConvArgs& a, const Plan& pl, const Tables& T, Cursor& c, int h) {
  if (c.have) {
    c.sl += 2;
    if (c.sl * 64 < c.s.n_valid) return true;
    c.sl = h;
    if (++c.mt < c.s.mt_cnt) return true;
    c.have = 0;
  }
  while (walker_next<M>(a, pl, T, c.w, c.s)) {          // next sub-item in which this half has a slab
    if (h * 64 < c.s.n_valid) {
      c.have = 1;
      c.mt = 0;
      c.sl = h;
      return true;
    }
  }
  return false;
}

// epilogue column tables of one sub-item: folded-BN scale / shift of tile column c and (OUT_ROWS) its compact position
template <class M>
__device__ __forceinline__ void column_entry(const ConvArgs& a, const Plan& pl, const Sub& s, int c, float& sc, float& sh,
                                             int& pos) {
  const int jj = s.n0 + c;
  int o = -1;
  pos = -1;
  if (c < s.n_valid) {
    if (M::omode(pl) == OUT_ROWS) {
      // real channel jj: active iff its group is in the sample's ascending list; its rank is the compact position
      if (a.n_idx) {
        const int grp = jj / a.n_gran, na = s.Nc / a.n_gran;
        const int* lst = a.n_idx + (size_t)s.b * a.n_ld;
        int lo = 0, hi = na;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(lst + mid) < grp) lo = mid + 1; else hi = mid;
        }
        if (lo < na && __ldg(lst + lo) == grp) { o = jj; pos = lo * a.n_gran + jj % a.n_gran; }
      } else {
        o = jj; pos = jj;
      }
    } else if (jj < s.Nc) {
      o = a.n_idx ? __ldg(a.n_idx + (size_t)s.b * a.n_ld + jj / a.n_gran) * a.n_gran + jj % a.n_gran : jj;
      pos = o;
    }
  }
  sc = o >= 0 ? 1.f : 0.f;                                       // inactive / pad columns come out as exact zeros
  sh = 0.f;
  if (o >= 0 && a.scale) { sc = __ldg(a.scale + o); sh = __ldg(a.shift + o); }
  if (o >= 0 && a.n_mask && __ldg(a.n_mask + (size_t)s.b * (a.C_out / a.n_mask_gran) + o / a.n_mask_gran) == 0) sc = 0.f;   // mask before BN
}

// Masked-dense channel gate of 32 consecutive output channels [ch0, ch0 + 32) of one sample: bit e set = channel ch0 + e
// is kept.  `mrow` = the sample's row of n_mask (one byte per group of `gran` channels).  fast16: gran == 2 and the 16
// bytes are aligned - one load.
__device__ __forceinline__ uint32_t nmask_bits(const uint8_t* mrow, int ch0, int gran, int c_out, bool fast16) {
  uint32_t keep = 0u;
  if (fast16) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(mrow + (ch0 >> 1)));
    const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((wv[i] >> (8 * j)) & 0xffu) keep |= 3u << (2 * (4 * i + j));
  } else {
    for (int e = 0; e < 32; ++e)
      if (ch0 + e < c_out && __ldg(mrow + (ch0 + e) / gran) != 0) keep |= 1u << e;
  }
  return keep;
}

// One 32-column pass of the slab epilogue for this thread's pixel row: accumulator (registers) -> folded BN ->
// gate -> (+ residual from the slab) -> fp16 -> ReLU -> back into the slab.  The flags are template parameters so the
// eight 16-byte groups are straight-line code: the table / residual loads of all groups issue back to back.
template <bool RES, bool RELU>
__device__ __forceinline__ void slab_pass(float* v, uint32_t t_scale, uint32_t t_shift, int c0, uint32_t srow, uint32_t sw,
                                          int p, float gate, uint32_t keep = 0xffffffffu) {
#pragma unroll
  for (int gp = 0; gp < 4; gp += 2) {             // two 16-byte groups at a time: loads first, then arithmetic
    float4 s0[2], s1[2], h0[2], h1[2];
    uint4 r4[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int g4 = gp + j;
      s0[j] = lds_f4(t_scale + (uint32_t)(c0 + g4 * 8) * 4u);
      s1[j] = lds_f4(t_scale + (uint32_t)(c0 + g4 * 8 + 4) * 4u);
      h0[j] = lds_f4(t_shift + (uint32_t)(c0 + g4 * 8) * 4u);
      h1[j] = lds_f4(t_shift + (uint32_t)(c0 + g4 * 8 + 4) * 4u);
      if (RES) r4[j] = lds128(srow + ((((uint32_t)(p * 4 + g4)) ^ sw) << 4));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int g4 = gp + j;
      float* w = v + g4 * 8;
      if (keep != 0xffffffffu) {                   // masked-dense channel gate: a gated channel's accumulator counts as 0
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (!((keep >> (g4 * 8 + e)) & 1u)) w[e] = 0.f;
      }
      w[0] = fmaf(w[0], s0[j].x, h0[j].x) * gate; w[1] = fmaf(w[1], s0[j].y, h0[j].y) * gate;
      w[2] = fmaf(w[2], s0[j].z, h0[j].z) * gate; w[3] = fmaf(w[3], s0[j].w, h0[j].w) * gate;
      w[4] = fmaf(w[4], s1[j].x, h1[j].x) * gate; w[5] = fmaf(w[5], s1[j].y, h1[j].y) * gate;
      w[6] = fmaf(w[6], s1[j].z, h1[j].z) * gate; w[7] = fmaf(w[7], s1[j].w, h1[j].w) * gate;
      if (RES) {
        const __half2* rh = reinterpret_cast<const __half2*>(&r4[j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(rh[e]);
          w[2 * e] += f.x;
          w[2 * e + 1] += f.y;
        }
      }
      uint4 o4;
      __half2* oh = reinterpret_cast<__half2*>(&o4);
#pragma unroll
      for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(w[2 * e], w[2 * e + 1]);
      if (RELU) {                                  // max(.,0) commutes with the rounding to fp16
        const __half2 z = __float2half2_rn(0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e) oh[e] = __hmax2(oh[e], z);
      }
      sts128(srow + ((((uint32_t)(p * 4 + g4)) ^ sw) << 4), o4);
    }
  }
}

// ------------------------------------------------------------------ the kernel
template <int SPEC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tma_kernel(const __grid_constant__ ConvArgs a, const __grid_constant__ Plan pl,
                const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_r) {
  using M = Mode<SPEC>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
  unsigned char* stg = smem + pl.a_region_bytes + (size_t)pl.stages * pl.stage_bytes;   // slab rings / row staging (1024-aligned)
  Tables& T = *reinterpret_cast<Tables*>(stg + pl.stg_bytes);
  const uint32_t a_base = smem_u32(smem);                           // halo mode: two activation slots in front of the stages
  const uint32_t smem_base = a_base + pl.a_region_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HWo = a.H_out * a.W_out;
  const int taps = a.ksize * a.ksize;

  pdl_launch_dependents();               // the next kernel's CTAs may take an SM as soon as one of ours leaves it
  if (threadIdx.x == 0) {
    for (int s = 0; s < pl.stages; ++s) {
      mbar_init(&T.full[s], pl.full_count);
      mbar_init(&T.empty[s], pl.dual ? 2 : 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&T.tfull[i], pl.dual ? 2 : 1);
      mbar_init(&T.tempty[i], EPI_WARPS);        // one arrival per epilogue warp (256 arrivals on one word serialise)
    }
    for (int h = 0; h < 2; ++h)
      for (int i = 0; i < MAX_RING; ++i) mbar_init(&T.rfull[h][i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&T.afull[i], 1);
      mbar_init(&T.aempty[i], pl.dual ? 2 : 1);
    }
    for (int h = 0; h < 2; ++h)
      for (int i = 0; i < MAX_RING; ++i) {
        mbar_init(&T.sready[h][i], EPI_WARPS / 2);
        mbar_init(&T.sfree[h][i], 1);
        mbar_init(&T.gdone[h][i], 2);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float* stab = reinterpret_cast<float*>(&T + 1);
  if (pl.static_cols)                                             // (weights: not produced by the previous kernel)
    for (int i = threadIdx.x; i < pl.stab_cols; i += NUM_THREADS) {
      const bool in = i < a.C_out;
      stab[i] = in ? (a.scale ? __ldg(a.scale + i) : 1.f) : 0.f;
      stab[pl.stab_cols + i] = (in && a.scale) ? __ldg(a.shift + i) : 0.f;
    }
  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&map_a);
    if (M::bmode(pl) == BMODE_TMA) tma_prefetch_desc(&map_b);
  }
  if (warp == 0 && lane == 0 && M::omode(pl) == OUT_SLAB) {
    tma_prefetch_desc(&map_y);
    if (M::has_res(a)) tma_prefetch_desc(&map_r);
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&T.tmem_base)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- everything above touches nothing an earlier kernel of the stream writes; everything below may
  pdl_wait();
  if (pl.cnt_cached)                                              // one round trip for every per-sample count
    for (int i = threadIdx.x; i < a.B; i += NUM_THREADS) {
      T.kcnt[i] = a.k_idx ? __ldg(a.k_cnt + i) : 0;
      T.ncnt[i] = a.n_idx ? __ldg(a.n_cnt + i) : 0;
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = T.tmem_base;
  Walker wk;
  walker_init(a, pl, wk);
  Sub s;

  if (warp == TMA_WARP) {
    // =========================================================== TMA producer
    if (M::bmode(pl) == BMODE_G4 && pl.g4_cp) {
      // ---- channel skipping, weight rows by cp.async (gather warps): this warp streams the activation tiles only -
      //      one per sub-item and 64-channel chunk, in order, as soon as one of the two slots is free
      if (lane == 0) {
        Walker wa;
        walker_init(a, pl, wa);
        Sub sa;
        int a_next = 0;
        while (walker_next<M>(a, pl, T, wa, sa))
          for (int kq = 0; kq < sa.cpt; ++kq, ++a_next) {
            const int slot = a_next & 1;
            mbar_wait(&T.aempty[slot], (uint32_t)((a_next >> 1) & 1) ^ 1u);
            mbar_arrive_expect_tx(&T.afull[slot], (uint32_t)pl.a_tx);
            tma_load_4d(a_base + slot * pl.a_slot_bytes, &map_a, &T.afull[slot], kq * 64, -1, (sa.mt0) * pl.R - 1, sa.b);
          }
      }
      __syncwarp();
    } else if (M::bmode(pl) == BMODE_G4) {
      // ---- channel skipping: the WHOLE warp produces.  Lane 0 stages the activation tiles (halo mode, as below); every
      //      lane issues the tile::gather4 copies of its four-row groups of the sample's ACTIVE weight rows: group j of
      //      a stage lands at byte j * 512 of the B tile - rows 4j .. 4j+3 of a K-major SWIZZLE_128B tile (the swizzle
      //      is a function of the shared-memory address, measured: scripts/mma_rate.cu, profiles/r02b_mma_rate.txt).
      int stage = 0;
      uint32_t phase = 0;
      Walker wa;
      walker_init(a, pl, wa);
      Sub sa;
      bool a_more = walker_next<M>(a, pl, T, wa, sa);
      int a_kq = 0, a_next = 0, g = 0;
      auto issue_a = [&]() {                                       // (all lanes keep the same cursor state)
        const int slot = a_next & 1;
        if (lane == 0) {
          mbar_arrive_expect_tx(&T.afull[slot], (uint32_t)pl.a_tx);
          tma_load_4d(a_base + slot * pl.a_slot_bytes, &map_a, &T.afull[slot], a_kq * 64, -1, (sa.mt0) * pl.R - 1, sa.b);
        }
        ++a_next;
        if (++a_kq >= sa.cpt) {
          a_kq = 0;
          a_more = walker_next<M>(a, pl, T, wa, sa);
        }
      };
      while (walker_next<M>(a, pl, T, wk, s)) {
        // this lane's weight rows for the sub-item: compact columns n0 + 4 (32 q + lane) + e, e < 4, q < 2
        const int real = max(0, min(s.Nc - s.n0, pl.BN));            // active columns of this tile
        const int ngath = (real + 3) >> 2;                           // gather4 copies per stage
        int rows[2][4];
#pragma unroll
        for (int q2 = 0; q2 < 2; ++q2)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int jj = s.n0 + 4 * (32 * q2 + lane) + e;
            rows[q2][e] = jj < s.Nc ? __ldg(a.n_idx + (size_t)s.b * a.n_ld + jj / a.n_gran) * a.n_gran + jj % a.n_gran : 0;
          }
        const uint32_t tx = (uint32_t)ngath * 512u;
        for (int kq = 0; kq < s.cpt; ++kq, ++g) {
          while (a_more && a_next <= g) {                            // the tile this chunk's MMAs read: must be on its way
            mbar_wait(&T.aempty[a_next & 1], (uint32_t)((a_next >> 1) & 1) ^ 1u);
            issue_a();
          }
          for (int tap = 0; tap < taps; ++tap) {
            if (a_more && a_next == g + 1) {                         // next chunk's tile as soon as its slot is free
              const uint32_t ok = mbar_try(smem_u32(&T.aempty[a_next & 1]), (uint32_t)((a_next >> 1) & 1) ^ 1u);
              if (__shfl_sync(0xffffffffu, ok, 0)) issue_a();
            }
            mbar_wait(&T.empty[stage], phase ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&T.full[stage], tx);
            __syncwarp();
            const uint32_t Bs = smem_base + stage * pl.stage_bytes;
            const int col = tap * a.C_in + kq * 64;
            if (lane < ngath) tma_gather4(Bs + (uint32_t)lane * 512u, &map_b, &T.full[stage], col, rows[0][0], rows[0][1], rows[0][2], rows[0][3]);
            if (32 + lane < ngath)
              tma_gather4(Bs + (uint32_t)(32 + lane) * 512u, &map_b, &T.full[stage], col, rows[1][0], rows[1][1], rows[1][2], rows[1][3]);
            if (++stage == pl.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      KP_DECL;
      if (M::halo(pl)) {
        // two streams issued by this one thread: activation tiles (one per sub-item and 64-channel chunk, kept one
        // chunk ahead, also across sub-items) and weight tiles (one per chunk and tap)
        Walker wa;
        walker_init(a, pl, wa);
        Sub sa;
        bool a_more = walker_next<M>(a, pl, T, wa, sa);
        while (a_more && sa.cpt == 0) a_more = walker_next<M>(a, pl, T, wa, sa);
        int a_kq = 0, a_next = 0, g = 0;
        // issue the next activation tile (chunk a_next of the A stream) into its slot
        auto issue_a = [&]() {
          const int slot = a_next & 1;
          if (pl.dbg & 2) mbar_arrive(&T.afull[slot]);
          else {
          mbar_arrive_expect_tx(&T.afull[slot], (uint32_t)pl.a_tx);
          tma_load_4d(a_base + slot * pl.a_slot_bytes, &map_a, &T.afull[slot], a_kq * 64, -1, (sa.mt0) * pl.R - 1, sa.b);
          }
          ++a_next;
          if (++a_kq >= sa.cpt) {
            a_kq = 0;
            a_more = walker_next<M>(a, pl, T, wa, sa);
            while (a_more && sa.cpt == 0) a_more = walker_next<M>(a, pl, T, wa, sa);
          }
        };
        while (walker_next<M>(a, pl, T, wk, s)) {
          for (int kq = 0; kq < s.cpt; ++kq, ++g) {
            while (a_more && a_next <= g) {                       // the tile this chunk's MMAs read: must be on its way
              mbar_wait(&T.aempty[a_next & 1], (uint32_t)((a_next >> 1) & 1) ^ 1u);
              issue_a();
            }
            KP_LAP(0);
            for (int tap = 0; tap < taps; ++tap) {
              // the NEXT chunk's tile goes out as soon as its slot is free - without holding up the weight stream
              if (a_more && a_next == g + 1 &&
                  mbar_try(smem_u32(&T.aempty[a_next & 1]), (uint32_t)((a_next >> 1) & 1) ^ 1u))
                issue_a();
              mbar_wait(&T.empty[stage], phase ^ 1);
              KP_LAP(1);
              mbar_arrive_expect_tx(&T.full[stage], (uint32_t)pl.b_tx);
              tma_load_2d(smem_base + stage * pl.stage_bytes, &map_b, &T.full[stage], tap * a.C_in + kq * 64, s.n0);
              if (++stage == pl.stages) { stage = 0; phase ^= 1; }
              KP_LAP(2);
            }
          }
        }
      } else
      while (walker_next<M>(a, pl, T, wk, s)) {
        KP_LAP(0);                                   // decode
        const uint32_t tx = (uint32_t)(((pl.dbg & 2) ? 0 : s.mt_cnt * pl.a_tx) + (M::bmode(pl) == BMODE_TMA ? pl.b_tx : 0));
        for (int tap = 0; tap < taps; ++tap) {
          const int ty = tap / a.ksize, tx_ = tap - ty * a.ksize;
          for (int kq = 0; kq < s.cpt; ++kq) {
            const int k0 = kq * 64;
            mbar_wait(&T.empty[stage], phase ^ 1);
            KP_LAP(1);                               // wait for a free stage
            const uint32_t As = smem_base + stage * pl.stage_bytes;
            mbar_arrive_expect_tx(&T.full[stage], tx);
            for (int m = 0; m < ((pl.dbg & 2) ? 0 : s.mt_cnt); ++m) {
              const int mt = s.mt0 + m;
              if (pl.R) tma_load_4d(As + m * A_TILE_BYTES, &map_a, &T.full[stage], k0, tx_ - a.pad, mt * pl.R * a.stride + ty - a.pad, s.b);
              else tma_load_3d(As + m * A_TILE_BYTES, &map_a, &T.full[stage], k0, mt * BM, s.b);
            }
            if (M::bmode(pl) == BMODE_TMA)
              tma_load_2d(As + pl.b_off, &map_b, &T.full[stage], tap * a.C_in + k0, s.n0);
            if (++stage == pl.stages) { stage = 0; phase ^= 1; }
            KP_LAP(2);                               // issue
          }
        }
        if (s.has_bias) {                            // the H1-constant step is staged by the gather warps only
          mbar_wait(&T.empty[stage], phase ^ 1);
          mbar_arrive(&T.full[stage]);
          if (++stage == pl.stages) { stage = 0; phase ^= 1; }
          KP_LAP(1);
        }
      }
      KP_FLUSH(0);
    }
    __syncwarp();
  } else if (warp == MMA_WARP || (pl.dual && warp == MMA_WARP2)) {
    // =========================================================== MMA issuer
    // The whole warp runs this loop with identical values and an elected lane issues (see umma_f16_elect).
    {
      const int m_first = (pl.dual && warp == MMA_WARP2) ? 1 : 0, m_step = pl.dual ? 2 : 1;   // this warp's m-tiles
      int stage = 0, buf = 0, hg = 0;
      uint32_t phase = 0, bphase = 0;
      KP_DECL;
      // Everything the MMA operands are built from is the same in all 32 lanes, but values that came through memory
      // (the TMEM base address, the shared-memory window, decoded work items) live in per-thread registers as far as the
      // compiler knows, and every tcgen05.mma then needs seven R2UR.BROADCAST moves into uniform registers.  A full-warp
      // shuffle from lane 0 tells the compiler the value is warp-uniform: descriptors are then computed on the uniform
      // datapath and an MMA costs a handful of issue slots - this warp shares its scheduler with two epilogue warps.
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t a_base_u = __shfl_sync(0xffffffffu, a_base, 0), smem_base_u = __shfl_sync(0xffffffffu, smem_base, 0);
      while (walker_next<M>(a, pl, T, wk, s)) {
        s.mt_cnt = __shfl_sync(0xffffffffu, s.mt_cnt, 0);
        s.cpt = __shfl_sync(0xffffffffu, s.cpt, 0);
        s.nk16 = __shfl_sync(0xffffffffu, s.nk16, 0);
        s.umma_n = __shfl_sync(0xffffffffu, s.umma_n, 0);
        s.nchunks = __shfl_sync(0xffffffffu, s.nchunks, 0);
        s.has_bias = __shfl_sync(0xffffffffu, s.has_bias, 0);
        KP_LAP(0);                                               // decode
        mbar_wait(&T.tempty[buf], bphase ^ 1);                   // epilogue has drained this buffer
        KP_LAP(1);                                               // wait for a free accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base_u + buf * (pl.MT * pl.acc_cols);
        const uint32_t idesc = umma_idesc_f16((pl.dbg & 16) ? s.umma_n / 2 : s.umma_n, M::bmode(pl) == BMODE_KROWS);   // (dbg 16: half-N MMAs, timing only)
        if (M::halo(pl)) {
          for (int kq = 0; kq < s.cpt; ++kq, ++hg) {
            const int aslot = hg & 1;
            const int n16 = min(4, s.nk16 - kq * 4);
            mbar_wait(&T.afull[aslot], (uint32_t)((hg >> 1) & 1));
            KP_LAP(2);
            tc_fence_after();
            const uint32_t Aslot = a_base_u + aslot * pl.a_slot_bytes;
            for (int tap = 0; tap < taps; ++tap) {
              const int ty = tap / 3, tx_ = tap - ty * 3;
              mbar_wait(&T.full[stage], phase);
              KP_LAP(2);
              if (M::bmode(pl) == BMODE_G4) fence_proxy_async();   // (cp.async rows: generic-proxy writes -> async proxy)
              tc_fence_after();
              const uint64_t bd = umma_desc(smem_base_u + stage * pl.stage_bytes, 16, 1024);
              if (!(pl.dbg & 4))
              for (int m = m_first; m < s.mt_cnt; m += m_step) {
                // tile m, tap (ty,tx): 128 consecutive rows of the padded image starting at row (m R + ty) Wp + tx
                const uint32_t aaddr = Aslot + (uint32_t)(((m * pl.R + ty) * pl.Wp + tx_) * 128);
                const uint64_t ad = umma_desc(aaddr, 16, 1024) | (pl.halo_bo ? ((uint64_t)((aaddr >> 7) & 7u) << 49) : 0ull);
                if (n16 == 4) umma_f16_elect_x4(d_tmem + m * pl.acc_cols, ad, bd, idesc, (kq | tap) ? 1u : 0u, 2u);
                else
                for (int k = 0; k < n16; ++k)
                  umma_f16_elect(d_tmem + m * pl.acc_cols, ad + 2 * k, bd + 2 * k, idesc, (kq | tap | k) ? 1u : 0u);
              }
              KP_LAP(3);
              umma_commit_elect(&T.empty[stage]);
              if (++stage == pl.stages) { stage = 0; phase ^= 1; }
              KP_LAP(5);
            }
            umma_commit_elect(&T.aempty[aslot]);                   // the activation slot is free once these MMAs retire
          }
        } else {
          int kq = 0;                                            // 64-channel chunk within the tap
          for (int ch = 0; ch < s.nchunks; ++ch) {
            const bool bias_step = s.has_bias && ch == s.nchunks - 1;
            const int n16 = bias_step ? 1 : min(4, s.nk16 - kq * 4);
            if (++kq == s.cpt) kq = 0;
            mbar_wait(&T.full[stage], phase);
            KP_LAP(2);                                           // wait for operands
            if (M::bmode(pl) != BMODE_TMA) fence_proxy_async();  // cp.async (generic proxy) writes -> async proxy
            tc_fence_after();
            KP_LAP(4);                                           // fences
            const uint32_t As = smem_base_u + stage * pl.stage_bytes;
            const uint32_t Bs = As + pl.b_off;
            const uint64_t bd = M::bmode(pl) == BMODE_KROWS ? umma_desc(Bs, 8192, 1024) : umma_desc(Bs, 16, 1024);
            const uint64_t bstep = M::bmode(pl) == BMODE_KROWS ? 128 : 2;
            if (!(pl.dbg & 4))
            for (int m = m_first; m < s.mt_cnt; m += m_step) {
              const uint64_t ad = umma_desc(As + m * A_TILE_BYTES, 16, 1024);
              if (n16 == 4) umma_f16_elect_x4(d_tmem + m * pl.acc_cols, ad, bd, idesc, ch ? 1u : 0u, (uint32_t)bstep);
              else
              for (int k = 0; k < n16; ++k)
                umma_f16_elect(d_tmem + m * pl.acc_cols, ad + 2 * k, bd + bstep * k, idesc, (ch | k) ? 1u : 0u);
            }
            KP_LAP(3);                                           // issue
            umma_commit_elect(&T.empty[stage]);                  // frees the stage when these MMAs retire
            if (++stage == pl.stages) { stage = 0; phase ^= 1; }
            KP_LAP(5);                                           // commit
          }
        }
        if (s.nchunks > 0) umma_commit_elect(&T.tfull[buf]);
        else if (lane == 0) mbar_arrive(&T.tfull[buf]);          // no active input channel: accumulator unused
        if (++buf == pl.nbuf) { buf = 0; bphase ^= 1; }
      }
      if (lane == 0 && warp == MMA_WARP) KP_FLUSH(1);
    }
    __syncwarp();
  } else if (warp >= GATHER_WARP0) {
    // =========================================================== weight gather (cp.async)
    if (M::bmode(pl) == BMODE_G4) {
      if (pl.g4_cp) {
        // the sample's ACTIVE weight rows (K-major w: one row = one output channel) -> rows of a swizzled K-major B tile;
        // thread = (16-byte chunk ac, rows ar0 + 24 i): the 8 threads of a row fetch its 128 contiguous bytes.
        // Stage order = the halo MMA loop's: 64-channel chunk outer, tap inner.
        const int pt = threadIdx.x - GATHER_WARP0 * 32;          // 0..191 (0..159 with a second MMA warp)
        const int ac = pt & 7, ar0 = pt >> 3;
        const int rstep = pl.dual ? 20 : 24;                     // rows ar0 + rstep * i
        // per thread, fixed for the whole kernel: where its rows' 16-byte chunk lands in a (swizzled) B tile
        uint32_t soff[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) soff[i] = sw128_off(ar0 + rstep * i, ac);
        int stage = 0;
        uint32_t phase = 0;
        KP_DECL;
        while (walker_next<M>(a, pl, T, wk, s)) {
          KP_LAP(0);
          // per sub-item: the weight row each of this thread's tile rows comes from (null: zero fill - pad rows)
          const __half* wrow[10];
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            const int row = ar0 + rstep * i, jj = s.n0 + row;
            wrow[i] = (row < s.umma_n && jj < s.Nc)
                          ? a.w + (size_t)(__ldg(a.n_idx + (size_t)s.b * a.n_ld + jj / a.n_gran) * a.n_gran + jj % a.n_gran) * taps * a.C_in + ac * 8
                          : nullptr;
          }
          const int nrow = (s.umma_n - ar0 + rstep - 1) / rstep;   // this thread's rows < umma_n (may be <= 0)
          for (int kq = 0; kq < s.cpt; ++kq) {
            const int n16 = min(4, s.nk16 - kq * 4);
            const bool kok = ac < 2 * n16 && kq * 64 + ac * 8 < a.C_in;
            KP_LAP(1);
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&T.empty[stage], phase ^ 1);
              KP_LAP(2);
              const uint32_t Bs = smem_base + stage * pl.stage_bytes;
              const int koff = tap * a.C_in + kq * 64;
              if (ac < 2 * n16) {
#pragma unroll
                for (int i = 0; i < 10; ++i)
                  if (i < nrow) {
                    const bool ok = kok && wrow[i] != nullptr;
                    cp_async_16(Bs + soff[i], ok ? wrow[i] + koff : a.w, ok ? 16u : 0u);
                  }
              }
              cp_async_arrive(&T.full[stage]);
              if (++stage == pl.stages) { stage = 0; phase ^= 1; }
              KP_LAP(3);
            }
          }
        }
        if (pt == 0) KP_FLUSH(2);
        asm volatile("cp.async.wait_all;" ::: "memory");
      }
    } else if (M::bmode(pl) == BMODE_ROWS || M::bmode(pl) == BMODE_KROWS) {
      const int pt = threadIdx.x - GATHER_WARP0 * 32;            // 0..191
      const int pw = pt >> 5;
      const int ac = pt & 7, ar0 = pt >> 3;                      // ROWS: 16-byte chunk, first row (rows ar0 + 24 i)
      int stage = 0;
      uint32_t phase = 0;
      int last_b = -1, last_mg = -1;
      KP_DECL;
      while (walker_next<M>(a, pl, T, wk, s)) {
        KP_LAP(0);                                               // decode
        const int cpr = s.umma_n >> 3;                           // 16-byte chunks per k-row (KROWS)
        int cpr2 = 2;
        while (cpr2 < cpr) cpr2 <<= 1;                           // lanes per k-row (power of two <= 32)
        const int rows_per_pass = 32 / cpr2 * GATHER_WARPS;
        const int kr0 = pw * (32 / cpr2) + lane / cpr2, kc = lane % cpr2;
        const bool kc_ok = kc < cpr && s.n0 + kc * 8 < a.C_out;
        const uint32_t kdst0 = (uint32_t)((kc >> 3) * 8192 + ((kc & 7) << 4));   // n-block + chunk (pre-swizzle)
        int browr[11];
        if (M::bmode(pl) == BMODE_ROWS) {
#pragma unroll
          for (int i = 0; i < 11; ++i) {
            const int row = ar0 + 24 * i, jj = s.n0 + row;
            browr[i] = (row < s.umma_n && jj < s.Nc)
                           ? (__ldg(a.n_idx + (size_t)s.b * a.n_ld + jj / a.n_gran) * a.n_gran + jj % a.n_gran) * taps * a.C_in
                           : -1;
          }
        } else if (s.b != last_b || (s.has_bias && s.mg != last_mg)) {
          // per-sample channel table / per-m-group tap-validity sets: rebuilt only when they change.  Every cp.async
          // that read the old tables was issued before these barriers (program order within each gather thread).
          named_bar_sync(1, GATHER_THREADS);
          if (s.b != last_b)
            for (int e = pt; e < s.Kc; e += GATHER_THREADS) {
              const int q = e / a.k_gran;
              T.kch[e] = __ldg(a.k_idx + (size_t)s.b * a.k_ld + q) * a.k_gran + (e - q * a.k_gran);
            }
          if (s.has_bias && s.mg != last_mg)
            for (int i = pt; i < s.mt_cnt * BM; i += GATHER_THREADS) {
              const int m = i / BM, rr = i & (BM - 1);
              int m0, rows;
              tile_rows(a, pl, s.mt0 + m, m0, rows);
              int vm = 0;
              if (rr < rows) {
                const int p = m0 + rr;
                const int oy = p / a.W_out, ox = p - oy * a.W_out;
                for (int tp = 0; tp < taps; ++tp) {
                  const int iy = oy * a.stride - a.pad + tp / a.ksize, ix = ox * a.stride - a.pad + tp % a.ksize;
                  if (iy >= 0 && iy < a.H_in && ix >= 0 && ix < a.W_in) vm |= 1 << tp;
                }
              }
              T.vmask[m][rr] = (unsigned short)vm;
            }
          named_bar_sync(1, GATHER_THREADS);
          last_b = s.b;
          last_mg = s.mg;
        }
        KP_LAP(1);                                               // index tables + barriers
        for (int tap = 0; tap < taps; ++tap) {
          const int tapk = tap * a.C_in;
          for (int kq = 0; kq < s.cpt; ++kq) {
            const int k0 = kq * 64;
            const int n16 = min(4, s.nk16 - kq * 4);
            mbar_wait(&T.empty[stage], phase ^ 1);
            KP_LAP(2);                                           // wait for a free stage
            const uint32_t Bs = smem_base + stage * pl.stage_bytes + pl.b_off;
            if (M::bmode(pl) == BMODE_ROWS) {
              if (ac < 2 * n16) {
                const int k = k0 + ac * 8;
                const bool kok = k < a.C_in;
                const __half* wk_ = a.w + tapk + k;
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                  const int row = ar0 + 24 * i;
                  if (row < s.umma_n) {
                    const bool ok = kok && browr[i] >= 0;
                    cp_async_16(Bs + sw128_off(row, ac), ok ? wk_ + browr[i] : a.w, ok ? 16u : 0u);
                  }
                }
              }
            } else if (kc < cpr) {
              const int nrows = 16 * n16;
              const __half* wn = a.wt + s.n0 + kc * 8;
              for (int kk = kr0; kk < nrows; kk += rows_per_pass) {
                const int e = k0 + kk;
                const bool ok = kc_ok && e < s.Kc;
                int rk = 0;
                if (ok) rk = T.kch[e];
                const uint32_t dst = Bs + (kdst0 ^ (uint32_t)((kk & 7) << 4)) + (kk >> 3) * 1024 + (kk & 7) * 128;
                cp_async_16(dst, ok ? wn + (size_t)(tapk + rk) * a.C_out : a.w, ok ? 16u : 0u);
              }
            }
            cp_async_arrive(&T.full[stage]);
            if (++stage == pl.stages) { stage = 0; phase ^= 1; }
            KP_LAP(3);                                           // issue
          }
        }
        if (s.has_bias) {
          // one more K=16 step: A' = tap-validity indicator of each pixel, B' = this sample's H1 constants
          // T[b, tap, o] (zero rows for tap >= taps): adds sum_{valid taps} T[b,tap,o] to the accumulator
          mbar_wait(&T.empty[stage], phase ^ 1);
          KP_LAP(2);
          const uint32_t As = smem_base + stage * pl.stage_bytes;
          const uint32_t Bs = As + pl.b_off;
          for (int i = pt; i < s.mt_cnt * BM * 2; i += GATHER_THREADS) {
            const int m = i / (BM * 2), rr = (i >> 1) & (BM - 1), c = i & 1;
            cp_async_16(As + m * A_TILE_BYTES + sw128_off(rr, c), g_vtab4 + (int)T.vmask[m][rr] * 16 + c * 8, 16u);
          }
          if (kc < cpr) {
            for (int kk = kr0; kk < 16; kk += rows_per_pass) {
              const bool ok = kc_ok && kk < taps;
              const uint32_t dst = Bs + (kdst0 ^ (uint32_t)((kk & 7) << 4)) + (kk >> 3) * 1024 + (kk & 7) * 128;
              cp_async_16(dst, ok ? a.bias_t + (size_t)s.b * a.bias_ld + (size_t)kk * a.C_out + s.n0 + kc * 8 : a.w,
                          ok ? 16u : 0u);
            }
          }
          cp_async_arrive(&T.full[stage]);
          if (++stage == pl.stages) { stage = 0; phase ^= 1; }
          KP_LAP(4);                                             // H1-constant step
        }
      }
      if (pt == 0) KP_FLUSH(2);
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (M::dma(pl) && warp < GATHER_WARP0 + 2 && lane == 0 && !(pl.dbg & 8)) {
      // =========================================================== slab DMA thread of epilogue half h
      // Issues the TMA store of every finished slab and keeps the residual slabs ahead of the epilogue, so the 128
      // epilogue threads of the half never wait for a copy to be ISSUED, only for data (rfull) or space (sfree).
      const int h = warp - GATHER_WARP0;
      unsigned char* ring = stg + (size_t)h * pl.ring * SLAB_BYTES;
      const bool has_res = M::has_res(a);
      Cursor cs, cl;
      walker_init(a, pl, cs.w);
      cs.mt = 0; cs.sl = 0; cs.have = 0;
      cl = cs;
      if (has_res)
        for (int t = 0; t < pl.ring; ++t) {
          if (!cursor_next<M>(a, pl, T, cl, h)) break;
          int m0, rows;
          tile_rows(a, pl, cl.s.mt0 + cl.mt, m0, rows);
          mbar_arrive_expect_tx(&T.rfull[h][t], (uint32_t)pl.r_tx);
          tma_load_3d(smem_u32(ring + t * SLAB_BYTES), &map_r, &T.rfull[h][t], cl.s.n0 + cl.sl * 64, m0, cl.s.b);
        }
      for (int j = 0; cursor_next<M>(a, pl, T, cs, h); ++j) {
        const int slot = j % pl.ring;
        int m0, rows;
        tile_rows(a, pl, cs.s.mt0 + cs.mt, m0, rows);
        mbar_wait(&T.sready[h][slot], (uint32_t)(j / pl.ring) & 1u);
        tma_store_3d(&map_y, smem_u32(ring + slot * SLAB_BYTES), cs.s.n0 + cs.sl * 64, m0, cs.s.b);
        bulk_commit();
        // wait until THIS store has been read out of its slab (a few hundred cycles; the next slab is ~2k cycles away)
        // and refill the slot at once: the residual of task j + ring is requested two task times before it is needed
        bulk_wait_read_n<0>();
        if (pl.gap) mbar_wait(&T.gdone[h][slot], (uint32_t)(j / pl.ring) & 1u);   // ... and pooled
        if (has_res) {
          if (cursor_next<M>(a, pl, T, cl, h)) {
            int pm0, prow;
            tile_rows(a, pl, cl.s.mt0 + cl.mt, pm0, prow);
            mbar_arrive_expect_tx(&T.rfull[h][slot], (uint32_t)pl.r_tx);
            tma_load_3d(smem_u32(ring + slot * SLAB_BYTES), &map_r, &T.rfull[h][slot], cl.s.n0 + cl.sl * 64, pm0, cl.s.b);
          }
        } else {
          mbar_arrive(&T.sfree[h][slot]);
        }
      }
      bulk_wait_all();
    } else if (pl.gap && warp >= GATHER_WARP0 + 2 && !(pl.dbg & 8)) {
      // =========================================================== pooling warps (fused GAP of the OUTPUT)
      // Two warps per epilogue half read every finished slab ([128 px][64 ch] fp16, exactly the values that go to
      // memory) once more and add up its columns per sample: lane = channel pair, each warp 64 of the 128 rows.
      // The per-(sample, tile) sums go to gap_partial[sample][k][C_out], k = tile - first tile of the sample; whoever
      // consumes them adds the (at most gap_tiles) partials of a sample in ascending k: a fixed order.
      const int gw = warp - (GATHER_WARP0 + 2);
      const int h = gw >> 1, part = gw & 1;
      unsigned char* ring = stg + (size_t)h * pl.ring * SLAB_BYTES;
      float* gp = stab + (pl.static_cols ? 2 * pl.stab_cols : 0) + h * (2 * 4 * 64);   // [task parity][half][part][segment][64 ch]
      const int hw = a.gap_hw;
      Cursor cs;
      walker_init(a, pl, cs.w);
      cs.mt = 0; cs.sl = 0; cs.have = 0;
      for (int j = 0; cursor_next<M>(a, pl, T, cs, h); ++j) {
        const int slot = j % pl.ring;
        int m0, rows;
        tile_rows(a, pl, cs.s.mt0 + cs.mt, m0, rows);            // flat layer: m0 = first pixel of the tile in [B*H*W]
        const int b_first = m0 / hw, b_last = (m0 + rows - 1) / hw;
        const int nseg = b_last - b_first + 1;                   // <= 4 (host checks hw)
        const uint32_t slab = smem_u32(ring + slot * SLAB_BYTES);
        const int r_lo = part * 64, r_hi = min(rows, part * 64 + 64);
        mbar_wait(&T.sready[h][slot], (uint32_t)(j / pl.ring) & 1u);
        float* gpj = gp + (j & 1) * (2 * 2 * 4 * 64);            // double-buffered by task parity: one barrier per task
        const uint32_t lane_off = ((uint32_t)lane & 3u) << 2, lane_chunk = (uint32_t)lane >> 2;
        for (int sg = 0; sg < nseg; ++sg) {
          const int lo = max(r_lo, (b_first + sg) * hw - m0), hi = min(r_hi, (b_first + sg + 1) * hw - m0);
          float acc[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] = 0.f;
          int r = lo;
#pragma unroll 2
          for (; r + 3 < hi; r += 4) {                           // four independent rows in flight per step
            uint32_t wv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              wv[u] = lds_u1(slab + (uint32_t)(r + u) * 128u + ((lane_chunk ^ ((uint32_t)(r + u) & 7u)) << 4) + lane_off);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wv[u]));
              acc[2 * u] += f.x;
              acc[2 * u + 1] += f.y;
            }
          }
          for (; r < hi; ++r) {
            const uint32_t w0 = lds_u1(slab + (uint32_t)r * 128u + ((lane_chunk ^ ((uint32_t)r & 7u)) << 4) + lane_off);
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w0));
            acc[0] += f.x;
            acc[1] += f.y;
          }
          *reinterpret_cast<float2*>(gpj + (part * 4 + sg) * 64 + 2 * lane) =
              make_float2((acc[0] + acc[2]) + (acc[4] + acc[6]), (acc[1] + acc[3]) + (acc[5] + acc[7]));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&T.gdone[h][slot]);           // the slab may be refilled
        named_bar_sync(6 + h, 64);                               // (also: everyone is done reading gp of task j - 1)
        {
          const int c = part * 32 + lane;
          const int tile = m0 / BM;
          for (int sg = 0; sg < nseg; ++sg) {
            const int bb = b_first + sg;
            const int k = tile - (bb * hw) / BM;
            a.gap_partial[((size_t)bb * a.gap_tiles + k) * a.C_out + cs.s.n0 + cs.sl * 64 + c] =
                gpj[sg * 64 + c] + gpj[(4 + sg) * 64 + c];
          }
        }
      }
    }
  } else {
    // =========================================================== epilogue
    const int et = threadIdx.x;                                  // 0..255 == tile column whose table entry this thread fills
    const int q = warp & 3, h = warp >> 2;                       // TMEM lane quadrant, column half
    const int row = q * 32 + lane;                               // accumulator row (TMEM lane) of this thread
    const bool elected = (warp == 4 * h) && lane == 0;           // issues this half's TMA copies
    unsigned char* ring = stg + (size_t)h * pl.ring * SLAB_BYTES;
    const bool relu_all = M::relu_mode(a) == LAUD_RELU_ALL;
    int buf = 0;
    uint32_t bphase = 0;
    int task = 0;                                                // slab tasks done by this half
    Cursor cur;                                                  // residual prefetcher (elected thread)
    walker_init(a, pl, cur.w);
    cur.mt = 0; cur.sl = 0; cur.have = 0;
    int pf = 0;                                                  // slab tasks whose residual load has been issued
    if (M::omode(pl) == OUT_SLAB && M::has_res(a) && elected && !M::dma(pl) && !(pl.dbg & 8)) {
      for (; pf < pl.ring - 1; ++pf) {
        if (!cursor_next<M>(a, pl, T, cur, h)) break;
        int m0, rows;
        tile_rows(a, pl, cur.s.mt0 + cur.mt, m0, rows);
        mbar_arrive_expect_tx(&T.rfull[h][pf % pl.ring], (uint32_t)pl.r_tx);
        tma_load_3d(smem_u32(ring + (pf % pl.ring) * SLAB_BYTES), &map_r, &T.rfull[h][pf % pl.ring],
                    cur.s.n0 + cur.sl * 64, m0, cur.s.b);
      }
    }
    KP_DECL;
    // column tables: entry `et` of the NEXT sub-item is fetched into registers while the current one is processed
    Sub nx;
    bool more = walker_next<M>(a, pl, T, wk, nx);
    int par = 0;
    const bool dyn_cols = !pl.static_cols;
    if (more && dyn_cols) {
      float sc, sh;
      int pos;
      column_entry<M>(a, pl, nx, et, sc, sh, pos);
      T.scale[0][et] = sc; T.shift[0][et] = sh; T.cpos[0][et] = pos;
    }
    int xb = -1;                                                 // OUT_EXPAND: sample whose expansion tables are built
    while (more) {
      s = nx;
      if (dyn_cols) named_bar_sync(2, EPI_THREADS);   // tables[par] are complete; nobody still reads tables[par ^ 1]
      if (M::omode(pl) == OUT_EXPAND && s.b != xb) {
        // dense word (channel pair) -> compact staging word, or the BN constant of a gated pair.  n_idx rows list every
        // group once (active ascending, then inactive); the previous sample's flush ended with a barrier of all epilogue
        // threads, and the flush that reads these tables starts with one.
        const int hw2 = a.n_gran >> 1, cnt = s.Nc / a.n_gran;
        for (int j = et; j < a.n_ld; j += EPI_THREADS) {
          const int grp = __ldg(a.n_idx + (size_t)s.b * a.n_ld + j);
          for (int t = 0; t < hw2; ++t) {
            const int w = grp * hw2 + t;
            float c0 = __ldg(a.shift + 2 * w), c1 = __ldg(a.shift + 2 * w + 1);
            if (relu_all) { c0 = fmaxf(c0, 0.f); c1 = fmaxf(c1, 0.f); }
            const __half2 hc = __floats2half2_rn(c0, c1);
            T.wsrc[w] = j < cnt ? (short)(j * hw2 + t) : (short)-1;
            T.cw[w] = *reinterpret_cast<const uint32_t*>(&hc);
          }
        }
        xb = s.b;
      }
      more = walker_next<M>(a, pl, T, wk, nx);
      float nsc = 0.f, nsh = 0.f;
      int npos = -1;
      if (more && dyn_cols) column_entry<M>(a, pl, nx, et, nsc, nsh, npos);
      KP_LAP(1);                                                 // decode + next table entry (loads in flight)
      // shared-space addresses of this sub-item's column tables
      const uint32_t t_scale = dyn_cols ? smem_u32(&T.scale[par][0]) : smem_u32(stab) + (uint32_t)s.n0 * 4u;
      const uint32_t t_shift = dyn_cols ? smem_u32(&T.shift[par][0]) : t_scale + (uint32_t)pl.stab_cols * 4u;
      const uint32_t t_cpos = smem_u32(&T.cpos[par][0]);
      mbar_wait(&T.tfull[buf], bphase);
      KP_LAP(2);                                                 // wait for the accumulator
      tc_fence_after();
      const uint32_t tbase = tmem_base + buf * (pl.MT * pl.acc_cols) + ((uint32_t)(q * 32) << 16);
      const bool have_acc = s.nchunks > 0;

      for (int m = 0; m < ((pl.dbg & 8) ? 0 : s.mt_cnt); ++m) {
        int m0, rows;
        tile_rows(a, pl, s.mt0 + m, m0, rows);
        // pixel of this accumulator row inside the tile (halo mode: rows are positions of the padded image)
        int prow = row;
        bool pvalid = row < rows;
        if (M::halo(pl)) {
          const int oyl = row / pl.Wp, ox = row - oyl * pl.Wp;
          prow = oyl * a.W_out + ox;
          pvalid = ox < a.W_out && prow < rows;
        }
        // masked-dense channel gate looked up per row (static column tables): this row's sample and its n_mask row
        const uint8_t* mrow = nullptr;
        if (!M::SLAB_FIXED && pl.nm && pvalid) {
          const int bb = pl.flat ? (m0 + prow) / a.gap_hw : s.b;
          mrow = a.n_mask + (size_t)bb * (a.C_out / a.n_mask_gran);
        }
        if (M::omode(pl) == OUT_SLAB) {
          bool row_on = true;                                    // spatial / layer gate of this pixel (one mask group)
          if (M::out_mask(a)) row_on = pvalid && M::out_mask(a)[(size_t)s.b * HWo + m0 + prow] != 0;
          for (int sl = h; sl * 64 < s.n_valid; sl += 2, ++task) {
            const int slot = task % pl.ring;
            unsigned char* slab = ring + slot * SLAB_BYTES;
            const uint32_t srow = smem_u32(slab) + (uint32_t)prow * 128u;
            const uint32_t sw = (uint32_t)(prow & 7);
            const bool has_res = M::has_res(a);
            if (has_res) mbar_wait(&T.rfull[h][slot], (uint32_t)(task / pl.ring) & 1u);
            else if (M::dma(pl) && task >= pl.ring) mbar_wait(&T.sfree[h][slot], (uint32_t)(task / pl.ring - 1) & 1u);
            KP_LAP(3);                                           // wait for the residual slab
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              const int c0 = sl * 64 + p * 32;
              if (c0 >= s.umma_n) break;
              float v[32];
              if (have_acc) {
                tmem_ld32(tbase + m * pl.acc_cols + c0, v);
              } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = 0.f;
              }
              if (!(M::halo(pl) && !pvalid)) {                       // (padding column of the padded image: not a pixel)
                // a gated-off VALID pixel has a finite accumulator, so multiplying by 0 zeroes it exactly;
                // RELU_WHERE_GATE0 keeps the value and applies the ReLU only where the gate is 0
                const bool gate0_relu = M::relu_mode(a) == LAUD_RELU_WHERE_GATE0;
                const float gate = (row_on || gate0_relu) ? 1.f : 0.f;
                const bool relu_row = relu_all || (gate0_relu && !row_on);
                uint32_t keep = 0xffffffffu;
                if (!M::SLAB_FIXED && mrow) keep = nmask_bits(mrow, s.n0 + c0, a.n_mask_gran, a.C_out, pl.nm_fast != 0);
                if (has_res) {
                  if (relu_row) slab_pass<true, true>(v, t_scale, t_shift, c0, srow, sw, p, gate, keep);
                  else slab_pass<true, false>(v, t_scale, t_shift, c0, srow, sw, p, gate, keep);
                } else {
                  if (relu_row) slab_pass<false, true>(v, t_scale, t_shift, c0, srow, sw, p, gate, keep);
                  else slab_pass<false, false>(v, t_scale, t_shift, c0, srow, sw, p, gate, keep);
                }
              }
            }
            KP_LAP(4);                                           // TMEM -> registers -> slab
            fence_proxy_async();                                 // generic-proxy slab writes -> visible to the TMA store
            if (M::dma(pl)) {
              __syncwarp();
              if (lane == 0) mbar_arrive(&T.sready[h][slot]);    // hand the slab to this half's DMA thread
              KP_LAP(5);
              continue;
            }
            if (elected && !M::has_res(a)) {                        // the NEXT task's slab must have left shared memory
              if (pl.ring == 3) bulk_wait_read_n<1>(); else bulk_wait_read_n<0>();
            }
            named_bar_sync(3 + h, HALF_THREADS);
            if (elected) {
              tma_store_3d(&map_y, smem_u32(slab), s.n0 + sl * 64, m0, s.b);
              bulk_commit();
              if (M::has_res(a)) {
                // slab of task-1 is free once its store has been read out; refill it for task + ring - 1
                bulk_wait_read_n<1>();
                if (pf == task + pl.ring - 1 && cursor_next<M>(a, pl, T, cur, h)) {
                  int pm0, prow;
                  tile_rows(a, pl, cur.s.mt0 + cur.mt, pm0, prow);
                  const int ps = pf % pl.ring;
                  mbar_arrive_expect_tx(&T.rfull[h][ps], (uint32_t)pl.r_tx);
                  tma_load_3d(smem_u32(ring + ps * SLAB_BYTES), &map_r, &T.rfull[h][ps], cur.s.n0 + cur.sl * 64, pm0,
                              cur.s.b);
                  ++pf;
                }
              }
            }
            KP_LAP(5);                                           // barrier + store + prefetch
          }
        } else if (M::omode(pl) == OUT_DIRECT) {
          // ---- OUT_DIRECT: long-K layers whose epilogue is a small share of the item: registers -> global,
          //      16 bytes per store, no staging (all shared memory goes to the operand pipeline)
          const bool valid = pvalid;
          bool row_on = true;
          if (a.out_mask) row_on = valid && a.out_mask[(size_t)s.b * HWo + m0 + prow] != 0;
          __half* yrow = a.y + ((size_t)s.b * HWo + m0 + prow) * a.ldy + s.n0;
          for (int c0 = h * 32; c0 < s.n_valid; c0 += 64) {
            float v[32];
            if (have_acc) {
              tmem_ld32(tbase + m * pl.acc_cols + c0, v);
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = 0.f;
            }
            if (valid) {
              uint4 o_prev = make_uint4(0u, 0u, 0u, 0u);
              uint32_t keep = 0xffffffffu;
              if (mrow) keep = nmask_bits(mrow, s.n0 + c0, a.n_mask_gran, a.C_out, pl.nm_fast != 0);
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                if (c0 + g4 * 8 < s.n_valid) {
                  const float4 s0 = lds_f4(t_scale + (uint32_t)(c0 + g4 * 8) * 4u);
                  const float4 s1 = lds_f4(t_scale + (uint32_t)(c0 + g4 * 8 + 4) * 4u);
                  const float4 h0 = lds_f4(t_shift + (uint32_t)(c0 + g4 * 8) * 4u);
                  const float4 h1 = lds_f4(t_shift + (uint32_t)(c0 + g4 * 8 + 4) * 4u);
                  float* w = v + g4 * 8;
                  if (keep != 0xffffffffu) {        // gated channel: accumulator counts as 0 (conv -> x mask -> bn)
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                      if (!((keep >> (g4 * 8 + e)) & 1u)) w[e] = 0.f;
                  }
                  w[0] = fmaf(w[0], s0.x, h0.x); w[1] = fmaf(w[1], s0.y, h0.y);
                  w[2] = fmaf(w[2], s0.z, h0.z); w[3] = fmaf(w[3], s0.w, h0.w);
                  w[4] = fmaf(w[4], s1.x, h1.x); w[5] = fmaf(w[5], s1.y, h1.y);
                  w[6] = fmaf(w[6], s1.z, h1.z); w[7] = fmaf(w[7], s1.w, h1.w);
                  if (!row_on && a.relu_mode != LAUD_RELU_WHERE_GATE0) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) w[e] = 0.f;
                  }
                  uint4 o4;
                  __half2* oh = reinterpret_cast<__half2*>(&o4);
#pragma unroll
                  for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(w[2 * e], w[2 * e + 1]);
                  if (relu_all || (a.relu_mode == LAUD_RELU_WHERE_GATE0 && !row_on)) {
                    const __half2 z = __float2half2_rn(0.f);
#pragma unroll
                    for (int e = 0; e < 4; ++e) oh[e] = __hmax2(oh[e], z);
                  }
                  // 32-byte stores where two pieces are complete and aligned: the scattered rows make this epilogue
                  // LSU-bound (one line per thread and instruction), so half the store instructions is half its time
                  if (pl.st256 && (g4 & 1) == 0 && c0 + g4 * 8 + 16 <= s.n_valid) {
                    o_prev = o4;
                  } else if (pl.st256 && (g4 & 1) == 1) {
                    stg256(yrow + c0 + (g4 - 1) * 8, o_prev, o4);
                  } else {
                    *reinterpret_cast<uint4*>(yrow + c0 + g4 * 8) = o4;
                  }
                }
              }
            }
          }
          KP_LAP(4);
        } else if (M::omode(pl) == OUT_EXPAND) {
          // ---- OUT_EXPAND: the accumulator columns are the sample's ACTIVE channels in ascending order.  Stage them
          //      compactly (fp16, folded BN + ReLU applied), then flush DENSE rows: every channel pair of the row comes
          //      from its compact word or, where gated, is the BN constant - coalesced 128-byte stores.
          const uint32_t srow = smem_u32(stg) + (uint32_t)prow * (uint32_t)pl.stg_pitch;
          for (int c0 = h * 32; c0 < s.n_valid; c0 += 64) {
            float v[32];
            tmem_ld32(tbase + m * pl.acc_cols + c0, v);
            if (pvalid) {
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                if (c0 + g4 * 8 < s.n_valid) {
                  const float4 s0 = lds_f4(t_scale + (uint32_t)(c0 + g4 * 8) * 4u);
                  const float4 s1 = lds_f4(t_scale + (uint32_t)(c0 + g4 * 8 + 4) * 4u);
                  const float4 h0 = lds_f4(t_shift + (uint32_t)(c0 + g4 * 8) * 4u);
                  const float4 h1 = lds_f4(t_shift + (uint32_t)(c0 + g4 * 8 + 4) * 4u);
                  float* w = v + g4 * 8;
                  w[0] = fmaf(w[0], s0.x, h0.x); w[1] = fmaf(w[1], s0.y, h0.y);
                  w[2] = fmaf(w[2], s0.z, h0.z); w[3] = fmaf(w[3], s0.w, h0.w);
                  w[4] = fmaf(w[4], s1.x, h1.x); w[5] = fmaf(w[5], s1.y, h1.y);
                  w[6] = fmaf(w[6], s1.z, h1.z); w[7] = fmaf(w[7], s1.w, h1.w);
                  uint4 o4;
                  __half2* oh = reinterpret_cast<__half2*>(&o4);
#pragma unroll
                  for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(w[2 * e], w[2 * e + 1]);
                  if (relu_all) {
                    const __half2 z = __float2half2_rn(0.f);
#pragma unroll
                    for (int e = 0; e < 4; ++e) oh[e] = __hmax2(oh[e], z);
                  }
                  sts128(srow + (uint32_t)(c0 + g4 * 8) * 2u, o4);          // column c of THIS n-tile
                }
              }
            }
          }
          KP_LAP(4);
          if (m == s.mt_cnt - 1) {                               // accumulators drained: the next sub-item's MMAs overlap the flush
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&T.tempty[buf]);
          }
          // flush this n-tile's share of the dense rows: the words whose compact position lies in [n0, n0 + n_valid) and,
          // with the sample's first n-tile, the BN constants of every gated pair
          named_bar_sync(5, EPI_THREADS);
          KP_LAP(6);                                             // (profiling: barrier after staging)
          {
            // Each lane owns the same four words (channel pairs) of every row: where they come from - a compact staging
            // word, the BN constant, or another n-tile's flush - is looked up ONCE; the row loop is then four independent
            // shared-memory loads and four coalesced stores (the straightforward per-word loop was a serial
            // LDS -> LDS -> STG chain: 12.7k cycles per tile, measured with the lap timers).
            const int nw = a.C_out >> 1, klo = s.n0 >> 1, khi = (s.n0 + s.n_valid) >> 1;
            const bool consts = s.n0 == 0;
            for (int w0 = lane; w0 < nw; w0 += 128) {
              int off[4];                                          // >= 0: byte offset in the staging row; -1: constant; -2: skip
              uint32_t cc[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int w = w0 + 32 * j;
                off[j] = -2;
                cc[j] = 0u;
                if (w < nw) {
                  const int k = T.wsrc[w];
                  if (k >= klo && k < khi) off[j] = (k - klo) * 4;
                  else if (k < 0 && consts) { off[j] = -1; cc[j] = T.cw[w]; }
                }
              }
              for (int r = warp; r < rows; r += EPI_WARPS) {
                const uint32_t src = smem_u32(stg) + (uint32_t)r * (uint32_t)pl.stg_pitch;
                uint32_t* dst = reinterpret_cast<uint32_t*>(a.y + ((size_t)s.b * HWo + m0 + r) * a.ldy) + w0;
                uint32_t v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = off[j] >= 0 ? lds_u1(src + (uint32_t)off[j]) : cc[j];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (off[j] >= -1) dst[32 * j] = v[j];
              }
            }
          }
          KP_LAP(7);                                             // (profiling: flush loop)
          named_bar_sync(5, EPI_THREADS);                        // staging may be overwritten by the next tile
          KP_LAP(5);
        } else {
          // ---- OUT_ROWS: compact the active real channels of this pixel row into the staging row
          const uint32_t srow = smem_u32(stg) + (uint32_t)row * (uint32_t)pl.stg_pitch;
          const bool row_ok = row < pl.stg_rows;
          for (int c0 = h * 32; c0 < s.n_valid; c0 += 64) {
            float v[32];
            if (have_acc) {
              tmem_ld32(tbase + m * pl.acc_cols + c0, v);
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = 0.f;
            }
            if (row_ok) {
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                const int pos = lds_i1(t_cpos + (uint32_t)(c0 + e) * 4u);   // channel pairs share a gate (even granularity)
                if (pos >= 0) {
                  float x0 = fmaf(v[e], lds_f1(t_scale + (uint32_t)(c0 + e) * 4u), lds_f1(t_shift + (uint32_t)(c0 + e) * 4u));
                  float x1 = fmaf(v[e + 1], lds_f1(t_scale + (uint32_t)(c0 + e + 1) * 4u), lds_f1(t_shift + (uint32_t)(c0 + e + 1) * 4u));
                  if (relu_all) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                  const int wd = pos >> 1;
                  const __half2 hv = __floats2half2_rn(x0, x1);
                  sts32(srow + (uint32_t)(((wd & ~31) | ((wd ^ row) & 31)) << 2), *reinterpret_cast<const uint32_t*>(&hv));
                }
              }
            }
          }
          KP_LAP(4);
          if (s.nt == pl.NT - 1) {
            // last n-tile of the row: zero pad [Nc, Nfill), then flush the tile's rows to y (coalesced words)
            if (h == 0 && row_ok)
              for (int j = s.Nc; j < s.Nfill; j += 2) {
                const int wd = j >> 1;
                sts32(srow + (uint32_t)(((wd & ~31) | ((wd ^ row) & 31)) << 2), 0u);
              }
            named_bar_sync(5, EPI_THREADS);
            const int nw = s.Nfill >> 1;
            for (int r = warp; r < rows; r += EPI_WARPS) {
              const uint32_t* src = reinterpret_cast<const uint32_t*>(stg + (size_t)r * pl.stg_pitch);
              uint32_t* dst = reinterpret_cast<uint32_t*>(a.y + ((size_t)s.b * HWo + m0 + r) * a.ldy);
              for (int w = lane; w < nw; w += 32) dst[w] = src[(w & ~31) | ((w ^ r) & 31)];
            }
            named_bar_sync(5, EPI_THREADS);                      // staging may be overwritten by the next tile
            KP_LAP(5);                                           // flush
          }
        }
      }
      KP_LAP(0);                                                 // loop exit
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && M::omode(pl) != OUT_EXPAND) mbar_arrive(&T.tempty[buf]);   // accumulators drained: the MMA warp may reuse them
                                                                 // (OUT_EXPAND arrived before its last flush)
      KP_LAP(7);                                                 // fence + arrive
      if (++buf == pl.nbuf) { buf = 0; bphase ^= 1; }
      par ^= 1;
      if (more && dyn_cols) { T.scale[par][et] = nsc; T.shift[par][et] = nsh; T.cpos[par][et] = npos; }
    }
    if (M::omode(pl) == OUT_SLAB && !M::dma(pl)) bulk_wait_all();
    KP_LAP(6);
    if (et == 0) KP_FLUSH(3);
    if (et == HALF_THREADS) KP_FLUSH(4);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 tensor map, 128-byte swizzle, innermost dimension first; strides in ELEMENTS for dims 1..rank-1
bool make_map(CUtensorMap* m, const void* base, int rank, const long long* dims, const long long* strides_elems,
              const int* box, const int* elem_strides = nullptr) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    es[i] = elem_strides ? (cuuint32_t)elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = (cuuint64_t)strides_elems[i - 1] * 2;
  }
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// one launch of an instantiation; `pdl`: programmatic dependent launch (prologue overlaps the predecessor's tail)
template <int SPEC>
void launch_conv_tma(int grid, size_t smem, cudaStream_t s, bool pdl, const ConvArgs& a, const Plan& pl, const CUtensorMap& ma,
                     const CUtensorMap& mb, const CUtensorMap& my, const CUtensorMap& mr) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, conv_tma_kernel<SPEC>, a, pl, ma, mb, my, mr);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Layouts the TMA kernel takes; everything else runs on the v3 kernel.
bool conv_tma_supported(const ConvArgs& a) {
  if (!conv_umma_supported(a)) return false;
  if (a.stride != 1 && a.stride != 2) return false;
  if (!((a.ksize == 1 && a.pad == 0) || (a.ksize == 3 && a.pad == 1))) return false;
  if (a.stride == 2 && (a.H_in != 2 * a.H_out || a.W_in != 2 * a.W_out || 2 * a.W_out > 256)) return false;
  if (a.row_idx || a.pre_bias) return false;
  if (a.out_mask && a.mask_groups != 1) return false;
  if (a.k_idx && !(a.wt && aligned16(a.wt))) return false;        // KUNITS layout
  if (a.k_idx && a.n_idx && (a.residual || a.out_mask || (a.n_gran & 1))) return false;
  if (a.n_idx && !a.k_idx && a.wt) return false;
  if (a.n_expand && !(a.n_idx && !a.k_idx && a.ksize == 3 && a.stride == 1 && a.pad == 1 && a.W_out + 2 <= BM && a.scale &&
                      !(a.n_gran & 1) && a.C_out <= 2 * XW_MAX && a.C_out % 8 == 0 && a.n_ld * a.n_gran == a.C_out &&
                      a.ldy >= a.C_out && (a.ldy & 1) == 0 && a.relu_mode != LAUD_RELU_WHERE_GATE0 && !a.residual && !a.out_mask))
    return false;
  if ((a.ksize == 3 || a.stride == 2) && a.W_out > BM) return false;
  if (a.ksize == 3 && a.stride == 1 && (a.H_in != a.H_out || a.W_in != a.W_out)) return false;
  if (a.k_idx && a.n_idx && round_up(a.C_out, 16) * 2 > 1024) return false;   // staging row
  if ((long long)a.B * a.H_out * a.W_out >= (1ll << 31)) return false;
  if (a.bias_t && !a.k_idx) return false;
  return encode_fn() != nullptr;
}

static int conv_forward_tma_x(const ConvArgs& a_in, cudaStream_t s, bool allow_row_gate);
// The per-row gate pays where a sample's pixels fill less than half an m-tile (7x7 maps: the flat GEMM packs 2.6 samples
// into a tile); measured on B200 (profiles/r02b_*): stage-4 conv1 -26 %, but +8 % on the 3x3 layers and +17 % on the 56x56
// slab layers, whose epilogues are issue-bound - those keep the per-item column tables.  LAUD_ROW_GATE=all|none overrides.
int conv_forward_tma(const ConvArgs& a_in, cudaStream_t s) {
  static const char* mode = getenv("LAUD_ROW_GATE");
  bool row_gate = a_in.ksize == 1 && a_in.stride == 1 && a_in.H_out * a_in.W_out <= 64;
  if (mode && !strcmp(mode, "all")) row_gate = true;
  if (mode && !strcmp(mode, "none")) row_gate = false;
  return conv_forward_tma_x(a_in, s, row_gate);
}

// allow_row_gate: the masked-dense channel gate (n_mask) may be looked up per accumulator row with static column
// tables (lets a gated 1x1 layer run flat); false = per-item column tables, per-sample items (the earlier scheme).
static int conv_forward_tma_x(const ConvArgs& a_in, cudaStream_t s, bool allow_row_gate) {
  // A 1x1 stride-1 layer with nothing per sample (no gather, gate or list) is one flat GEMM over all B*H*W pixels:
  // m-tiles run across sample boundaries, so images smaller than a tile (14x14, 7x7) leave no padded rows.
  ConvArgs a = a_in;
  bool flat = false;
  static const bool no_flat = getenv("LAUD_NO_FLAT") != nullptr;
  static const bool no_row_gate = getenv("LAUD_NO_ROW_GATE") != nullptr;        // A/B switch
  if (no_row_gate) allow_row_gate = false;
  if (!no_flat && a.ksize == 1 && a.stride == 1 && !a.k_idx && !a.n_idx && (!a.n_mask || allow_row_gate) && !a.sample_idx &&
      !a.out_mask && !a.bias_t && (long long)a.B * a.H_out * a.W_out < (1ll << 31)) {
    a.gap_hw = a.H_out * a.W_out;
    a.W_in = a.W_out = a.B * a.H_out * a.W_out;
    a.H_in = a.H_out = 1;
    a.B = 1;
    flat = true;
  }
  static int num_sms_dev[MAX_DEVICES] = {0};
  const int dev = current_device();
  if (num_sms_dev[dev] == 0) {
    int n = 0;
    LAUD_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    LAUD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    LAUD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    LAUD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    LAUD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    LAUD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    LAUD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    LAUD_CUDA(cudaFuncSetAttribute(conv_tma_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    vtab4_init_kernel<<<32, 256, 0, s>>>();
    if (int e = check_launch("vtab4_init_kernel")) return e;
    if (int e = finish_first_call_init(s, "conv_forward_tma")) return e;
    num_sms_dev[dev] = n;
  }
  const int num_sms = num_sms_dev[dev];
  Plan pl{};
  const int HWo = a.H_out * a.W_out;
  const int taps = a.ksize * a.ksize;
  pl.bmode = a.k_idx ? BMODE_KROWS : (a.n_idx ? (a.n_expand ? BMODE_G4 : BMODE_ROWS) : BMODE_TMA);
  const long long ktotal = (long long)taps * a.C_in;
  pl.omode = a.n_expand ? OUT_EXPAND : ((a.k_idx && a.n_idx) ? OUT_ROWS : ((!a.residual && ktotal >= 512) ? OUT_DIRECT : OUT_SLAB));
  const bool row_tiles = a.ksize == 3 || a.stride == 2;   // m-tiles of whole output rows: 4-d boxes (im2col / subsampling by TMA)
  static const bool no_halo = getenv("LAUD_NO_HALO") != nullptr;
  pl.halo = ((!no_halo || pl.bmode == BMODE_G4) && a.ksize == 3 && a.stride == 1 && (pl.bmode == BMODE_TMA || pl.bmode == BMODE_G4) &&
             !a.bias_t && a.W_out + 2 <= BM) ? 1 : 0;
  pl.Wp = a.W_out + 2;
  {
    const char* e = getenv("LAUD_HALO_BO");
    pl.halo_bo = e ? atoi(e) : 0;   // measured: the swizzle is a function of the absolute shared-memory address, a row-shifted start needs no base offset
  }
  if (row_tiles) {
    pl.R = BM / (pl.halo ? pl.Wp : a.W_out);
    if (pl.R > a.H_out) pl.R = a.H_out;
    // prefer an even split of the rows over one full and one nearly empty tile
    const int nt0 = (a.H_out + pl.R - 1) / pl.R;
    pl.R = (a.H_out + nt0 - 1) / nt0;
    pl.rows_per_tile = pl.R * a.W_out;
    pl.n_mtiles = (a.H_out + pl.R - 1) / pl.R;
  } else {
    pl.R = 0;
    pl.rows_per_tile = BM;
    pl.n_mtiles = (HWo + BM - 1) / BM;
  }
  pl.MT = pl.n_mtiles >= 2 ? 2 : 1;
  const int nfill_max = round_up(a.C_out, a.n_pad_align);
  const int span = pl.bmode == BMODE_KROWS ? a.C_out : nfill_max;
  if (pl.omode == OUT_EXPAND) {
    // compact columns in tiles of G4_BN: a sample's active count decides how many of its n-tiles exist (typically one);
    // all n-tiles of an m-group run inside ONE item, so consecutive sub-items share the activation tiles in L2
    pl.BN = span <= G4_BN ? round_up(span, 16) : G4_BN;
    pl.NT = (span + pl.BN - 1) / pl.BN;
    pl.NTI = pl.NT;                            // (every n-tile flushes its own columns: MT stays 2)
  } else if (pl.omode == OUT_ROWS) {
    pl.BN = span <= BN_MAX ? round_up(span, 16) : BN_MAX;
    pl.NT = (span + pl.BN - 1) / pl.BN;
    pl.NTI = pl.NT;
    if (pl.NTI > 1) pl.MT = 1;                 // one staging tile: all n-tiles of ONE m-tile before the flush
  } else {
    // wide tiles halve the re-reads of the activations when the reduction is long enough to hide a
    // non-overlapped epilogue (BN = 256 with two m-tiles fills TMEM: one accumulator buffer)
    // (an SS-mode tcgen05.mma costs >= ~128 cycles whatever its N - measured - so N = 256 is the only full-rate shape)
    static const bool slab256 = getenv("LAUD_SLAB_BN128") == nullptr;
    pl.BN = span <= 64 ? 64 : ((span > 128 && (pl.omode == OUT_DIRECT || slab256)) ? 256 : 128);
    if (pl.BN == 256 && pl.omode == OUT_SLAB) pl.MT = 1;   // keep two accumulator buffers: this epilogue must overlap the MMAs
    {
      static const int mt1 = getenv("LAUD_MT1") ? atoi(getenv("LAUD_MT1")) : 0;   // experiment: 1 -> 1x1 DIRECT layers, 2 -> 3x3 DIRECT layers
      if (pl.BN == 256 && pl.omode == OUT_DIRECT && ((a.ksize == 1 && (mt1 & 1)) || (a.ksize == 3 && (mt1 & 2)))) pl.MT = 1;
    }
    pl.NT = (span + pl.BN - 1) / pl.BN;
    pl.NTI = 1;
  }
  pl.n_mgroups = (pl.n_mtiles + pl.MT - 1) / pl.MT;
  const long long total = (long long)a.B * pl.n_mgroups * (pl.NT / pl.NTI);
  if (total >= (1ll << 31)) {
    set_error("conv_forward_tma: too many tiles");
    return LAUD_E_BADARG;
  }
  pl.total_items = (int)total;
  pl.acc_cols = pl.BN <= 64 ? 64 : (pl.BN <= 128 ? 128 : 256);
  pl.nbuf = (int)TMEM_COLS / (pl.MT * pl.acc_cols);
  if (pl.nbuf > 4) pl.nbuf = 4;
  const int b_bytes = pl.bmode == BMODE_KROWS ? ((pl.BN + 63) / 64) * 8192 : round_up(pl.BN, 4) * 128;
  pl.b_off = pl.halo ? 0 : pl.MT * A_TILE_BYTES;
  pl.stage_bytes = pl.b_off + round_up(b_bytes, 1024);
  int halo_box_rows = 0;
  if (pl.halo) {
    halo_box_rows = pl.MT * pl.R + 2;
    const int rows_read = ((pl.MT - 1) * pl.R + 2) * pl.Wp + 2 + BM;      // last row any tap of any tile may read (+1)
    const int rows_alloc = halo_box_rows * pl.Wp > rows_read ? halo_box_rows * pl.Wp : rows_read;
    pl.a_slot_bytes = round_up(rows_alloc * 128, 1024);
    pl.a_region_bytes = 2 * pl.a_slot_bytes;
    if (halo_box_rows > 256) pl.halo = 0;
  }
  if (!pl.halo) { pl.a_slot_bytes = 0; pl.a_region_bytes = 0; pl.b_off = pl.MT * A_TILE_BYTES; pl.stage_bytes = pl.b_off + round_up(b_bytes, 1024); }
  pl.ring = a.residual ? 3 : 2;
  pl.stg_pitch = round_up(round_up(nfill_max, 16) * 2, 128);
  pl.stg_rows = pl.rows_per_tile <= 64 ? 64 : BM;            // small images (7x7): half-height staging
  if (HWo < pl.stg_rows) pl.stg_rows = round_up(HWo, 32);
  if (pl.omode == OUT_EXPAND) {
    if (!pl.halo) {
      set_error("conv_forward_tma: n_expand needs the halo layout (3x3 stride 1, W_out + 2 <= 128)");
      return LAUD_E_UNSUPPORTED;
    }
    pl.stg_pitch = pl.BN * 2 + 16;                           // one n-tile's fp16 columns + 16 bytes: the 16-byte row stores
    pl.stg_rows = round_up(pl.rows_per_tile, 8);             // of a warp's 32 lanes fall on distinct banks
  }
  pl.cnt_cached = a.B <= CNT_CACHE ? 1 : 0;
  const int stg_bytes = pl.omode == OUT_SLAB ? 2 * pl.ring * SLAB_BYTES
                                             : ((pl.omode == OUT_ROWS || pl.omode == OUT_EXPAND) ? round_up(pl.stg_rows * pl.stg_pitch, 1024) : 0);
  pl.stg_bytes = stg_bytes;
  static const bool no_dma = getenv("LAUD_NO_DMA") != nullptr;
  pl.dma = (!no_dma && pl.omode == OUT_SLAB && pl.bmode == BMODE_TMA) ? 1 : 0;
  pl.NG = pl.NT / pl.NTI;
  {
    static const bool no256 = getenv("LAUD_NO_ST256") != nullptr;
    pl.st256 = (!no256 && (reinterpret_cast<uintptr_t>(a.y) & 31) == 0 && a.ldy % 16 == 0 && a.C_out % 16 == 0 && pl.BN % 16 == 0) ? 1 : 0;
  }
  // fused GAP of the output: flat layers whose slabs go through the DMA threads; a tile may span at most 4 samples
  pl.gap = 0;
  if (a.gap_partial) {
    const bool can = flat && pl.omode == OUT_SLAB && pl.dma && a.C_out % 64 == 0 && a.gap_hw >= 43 &&
                     a.gap_tiles >= (a.gap_hw - 1) / BM + 2;
    if (!can) {
      set_error("conv_forward_tma: fused GAP needs a flat 1x1 stride-1 layer without per-sample operands, C_out %% 64 == 0, "
                "H*W >= 43 and gap_tiles >= (H*W - 1) / 128 + 2");
      return LAUD_E_UNSUPPORTED;
    }
    pl.gap = 1;
  }
  const int gap_bytes = pl.gap ? 2 * 2 * 2 * 4 * 64 * 4 : 0;
  pl.simple = (!a.k_idx && !a.n_idx && !a.bias_t && pl.bmode == BMODE_TMA) ? 1 : 0;
  {
    static const int dbg = getenv("LAUD_DBG") ? atoi(getenv("LAUD_DBG")) : 0;
    pl.dbg = dbg;
  }
  pl.c_nfill = nfill_max;
  pl.c_nk16 = (a.C_in + 15) >> 4;
  pl.c_cpt = (pl.c_nk16 + 3) >> 2;
  pl.c_nchunks = pl.c_cpt * taps;
  pl.flat = flat ? 1 : 0;
  pl.static_cols = (pl.omode != OUT_ROWS && pl.omode != OUT_EXPAND && !a.n_idx && (!a.n_mask || allow_row_gate)) ? 1 : 0;
  pl.stab_cols = round_up(a.C_out, 64) + BN_MAX;              // reads of a partial last tile stay inside the (zero) padding
  int stab_bytes = pl.static_cols ? 2 * pl.stab_cols * 4 : 0;
  int avail = SMEM_LIMIT - 1024 - (int)sizeof(Tables) - stg_bytes - pl.a_region_bytes - stab_bytes - gap_bytes;
  if (pl.static_cols && avail / pl.stage_bytes < 2) {         // no room beside a two-stage pipeline: per-item tables
    if (a.n_mask) return conv_forward_tma_x(a_in, s, false);  // (the row gate needs the static tables)
    pl.static_cols = 0;
    avail += stab_bytes;
    stab_bytes = 0;
  }
  pl.nm = (a.n_mask && pl.static_cols) ? 1 : 0;
  pl.nm_fast = (pl.nm && a.n_mask_gran == 2 && a.C_out % 32 == 0 && (reinterpret_cast<uintptr_t>(a.n_mask) & 15) == 0) ? 1 : 0;
  {
    static const bool no_tsplit = getenv("LAUD_NO_TSPLIT") != nullptr;         // A/B switch
    pl.tsplit = (!no_tsplit && flat && pl.NG == 1 && pl.NTI == 1 && pl.MT == 2 && pl.omode == OUT_DIRECT && !pl.gap) ? 1 : 0;
  }
  pl.stages = avail / pl.stage_bytes;
  if (pl.stages > MAX_STAGES) pl.stages = MAX_STAGES;
  {
    static const int cap = getenv("LAUD_MAX_STAGES") ? atoi(getenv("LAUD_MAX_STAGES")) : 0;   // experiment: pipeline-depth sensitivity
    if (cap >= 2 && pl.stages > cap) pl.stages = cap;
  }
  if (pl.stages < 2) {
    if (pl.gap) { set_error("conv_forward_tma: no shared memory left for the fused GAP"); return LAUD_E_UNSUPPORTED; }
    if (pl.omode == OUT_EXPAND) { set_error("conv_forward_tma: no shared memory left for the n_expand pipeline"); return LAUD_E_UNSUPPORTED; }
    return conv_forward_umma(a, s);
  }
  {
    static const char* g4 = getenv("LAUD_G4");                  // "tma": stage the active weight rows by TMA gather4 (A/B switch)
    pl.g4_cp = (pl.bmode == BMODE_G4 && !(g4 && !strcmp(g4, "tma"))) ? 1 : 0;
  }
  {
    static const bool no_dual = getenv("LAUD_NO_DUAL_MMA") != nullptr;            // A/B switch
    // (slab layers too, unless warp 15 is a pooling warp of the fused GAP)
    pl.dual = (!no_dual && pl.MT == 2 && (pl.omode == OUT_DIRECT || pl.omode == OUT_EXPAND || (pl.omode == OUT_SLAB && !pl.gap)) &&
               (pl.bmode == BMODE_TMA || pl.bmode == BMODE_G4)) ? 1 : 0;
  }
  pl.full_count = pl.bmode == BMODE_G4 ? (pl.g4_cp ? (pl.dual ? GATHER_THREADS - 32 : GATHER_THREADS) : 1)
                                       : 1 + (pl.bmode == BMODE_TMA ? 0 : GATHER_THREADS);
  const int rows_box = pl.rows_per_tile < HWo ? pl.rows_per_tile : HWo;     // boxes never exceed the tensor extent
  int bn_box = pl.BN < a.C_out ? pl.BN : a.C_out;
  static const bool dbg_halfb = getenv("LAUD_DBG") && (atoi(getenv("LAUD_DBG")) & 1);   // TIMING EXPERIMENT ONLY (wrong results): half of every weight tile
  if (dbg_halfb && pl.bmode == BMODE_TMA) bn_box /= 2;
  pl.a_tx = pl.halo ? 128 * pl.Wp * halo_box_rows : 128 * rows_box;
  pl.b_tx = 128 * bn_box;
  pl.r_tx = 128 * rows_box;

  CUtensorMap map_a, map_b, map_y, map_r;
  memset(&map_b, 0, sizeof(map_b));
  memset(&map_y, 0, sizeof(map_y));
  memset(&map_r, 0, sizeof(map_r));
  bool ok = true;
  const long long cext_a = a.k_idx ? a.ldx : a.C_in;     // compact inputs may be read up to their pitch (zero padded by the producer)
  if (row_tiles) {
    // traversal stride = conv stride: the box spans stride*W_out x stride*R input pixels and delivers W_out x R of them
    const long long dims[4] = {cext_a, a.W_in, a.H_in, a.B};
    const long long str[3] = {a.ldx, (long long)a.W_in * a.ldx, (long long)a.H_in * a.W_in * a.ldx};
    const int box[4] = {64, pl.halo ? pl.Wp : a.W_out * a.stride, pl.halo ? halo_box_rows : pl.R * a.stride, 1};
    const int es[4] = {1, a.stride, a.stride, 1};
    ok = ok && make_map(&map_a, a.x, 4, dims, str, box, es);
  } else {
    const long long dims[3] = {cext_a, HWo, a.B};
    const long long str[2] = {a.ldx, (long long)HWo * a.ldx};
    const int box[3] = {64, rows_box, 1};
    ok = ok && make_map(&map_a, a.x, 3, dims, str, box);
  }
  if (pl.bmode == BMODE_TMA) {
    const long long dims[2] = {(long long)taps * a.C_in, a.C_out};
    const long long str[1] = {(long long)taps * a.C_in};
    const int box[2] = {64, bn_box};
    ok = ok && make_map(&map_b, a.w, 2, dims, str, box);
  } else if (pl.bmode == BMODE_G4) {            // tile::gather4: 2-d map, box {64 columns, 1 row}; four rows per copy
    const long long dims[2] = {(long long)taps * a.C_in, a.C_out};
    const long long str[1] = {(long long)taps * a.C_in};
    const int box[2] = {64, 1};
    ok = ok && make_map(&map_b, a.w, 2, dims, str, box);
  }
  if (pl.omode == OUT_SLAB) {
    // compact outputs (n_idx) may be overwritten up to their pitch (scratch, see laud_conv_desc); a dense output is clipped
    // at C_out, so a convolution can write a channel SLICE of a wider tensor (grouped convolutions run group by group)
    const long long dims[3] = {a.n_idx ? a.ldy : a.C_out, HWo, a.B};
    const long long str[2] = {a.ldy, (long long)HWo * a.ldy};
    const int box[3] = {64, rows_box, 1};
    ok = ok && make_map(&map_y, a.y, 3, dims, str, box);
    if (a.residual) {
      const long long rdims[3] = {a.C_out, HWo, a.B};
      const long long rstr[2] = {a.ldr, (long long)HWo * a.ldr};
      ok = ok && make_map(&map_r, a.residual, 3, rdims, rstr, box);
    }
  }
  if (!ok) {                                     // the driver refused a descriptor: v3 takes every layout v4 does
    if (pl.gap || pl.omode == OUT_EXPAND) { set_error("conv_forward_tma: tensor map refused"); return LAUD_E_UNSUPPORTED; }
    return conv_forward_umma(a, s);
  }

  const size_t smem = 1024 + (size_t)pl.a_region_bytes + (size_t)pl.stages * pl.stage_bytes + stg_bytes + sizeof(Tables) + stab_bytes + gap_bytes;
  const long long units = pl.tsplit ? pl.n_mtiles : total;                    // what the CTAs partition
  const int grid = (int)(units < num_sms ? units : num_sms);
  g_conv_paths[0].fetch_add(1, std::memory_order_relaxed);
  g_conv_tma_launches.fetch_add(1, std::memory_order_relaxed);
  {
    ConvProfScope prof(s);
    static const bool no_spec = getenv("LAUD_NO_SPEC") != nullptr;      // A/B switch: always the generic kernel
    int spec = 0;
    if (!no_spec && pl.simple && pl.bmode == BMODE_TMA) {
      if (pl.omode == OUT_SLAB && pl.dma && !pl.halo) {
        spec = 1;
        if (!a.out_mask && !a.n_mask && a.residual && a.relu_mode == LAUD_RELU_ALL) spec = 4;
        else if (!a.out_mask && !a.n_mask && !a.residual && a.relu_mode == LAUD_RELU_NONE) spec = 5;
      }
      else if (pl.omode == OUT_DIRECT) spec = pl.halo ? 3 : 2;
    }
    if (pl.bmode == BMODE_G4) spec = 6;
    // programmatic dependent launch: measured -2.5 % with the two graph chains of a large batch (early CTAs of one chain sit on
    // SMs the other chain could use), +0.8 % with one chain at batch 256, +4 % at batch 8 (configs[0]: 1.283 -> 1.234 ms) - the
    // engines switch it on for single-chain forwards of small batches (laud_conv_set_pdl); LAUD_PDL forces it on
    static const bool pdl_env = getenv("LAUD_PDL") != nullptr;
    const bool pdl = pdl_env || g_conv_pdl.load(std::memory_order_relaxed) != 0;
    switch (spec) {
      case 1: launch_conv_tma<1>(grid, smem, s, pdl, a, pl, map_a, map_b, map_y, map_r); break;
      case 2: launch_conv_tma<2>(grid, smem, s, pdl, a, pl, map_a, map_b, map_y, map_r); break;
      case 3: launch_conv_tma<3>(grid, smem, s, pdl, a, pl, map_a, map_b, map_y, map_r); break;
      case 4: launch_conv_tma<4>(grid, smem, s, pdl, a, pl, map_a, map_b, map_y, map_r); break;
      case 5: launch_conv_tma<5>(grid, smem, s, pdl, a, pl, map_a, map_b, map_y, map_r); break;
      case 6: launch_conv_tma<6>(grid, smem, s, pdl, a, pl, map_a, map_b, map_y, map_r); break;
      default: launch_conv_tma<0>(grid, smem, s, pdl, a, pl, map_a, map_b, map_y, map_r); break;
    }
  }
  return check_launch("conv_tma_kernel");
}

#ifdef LAUD_KPROF
// debug build only: copy (or reset) the lap timers of the last launches
extern "C" int laud_debug_kprof(long long* host_out /* [160][5][8] */, int reset) {
  void* p = nullptr;
  if (cudaGetSymbolAddress(&p, g_kprof) != cudaSuccess) return -1;
  const size_t n = sizeof(long long) * 160 * KP_ROLES * KP_SITES;
  if (reset) return cudaMemset(p, 0, n) == cudaSuccess ? 0 : -1;
  return cudaMemcpy(host_out, p, n, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
#endif

}  // namespace laud

extern "C" void laud_conv_set_pdl(int enable) { laud::g_conv_pdl.store(enable ? 1 : 0, std::memory_order_relaxed); }
