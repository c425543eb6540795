// The product kernel of the mask-conditioned convolution: a persistent,
// warp-specialised gather-GEMM on the 5th-generation tensor cores (tcgen05).
//
//   D[m, n] = sum_{tap, k} A[m, (tap,k)] * W[n, (tap,k)]
//     m : up to 128 output pixels of ONE sample (or of a device-side row list)
//     n : that sample's ACTIVE output channels (compact), <= 256 per tile
//     k : that sample's ACTIVE input channels (compact), per filter tap
//
// Roles (288 threads, one CTA per SM, grid = #SMs, static round-robin over
// work items (sample, m-tile, n-tile)):
//   warps 0-3  epilogue : tcgen05.ld the fp32 accumulator out of TMEM, apply the
//                         H1 pre-bias / folded BN / spatial gate / residual /
//                         ReLU, store fp16 (compact channels, zero pad);
//   warp  4    MMA      : one elected lane issues tcgen05.mma (M=128, N = the
//                         tile's runtime width, K=16) on 128B-swizzled K-major
//                         shared-memory operands; accumulators double-buffered
//                         in TMEM (2 x 256 columns) so the epilogue of tile i
//                         overlaps the main loop of tile i+1;
//   warps 5-8  producers: stage ONLY the active patches/channels: the im2col
//                         gather of the activation rows (zero-fill for padding
//                         and ragged tiles) and the per-sample row (n) and
//                         in-row (k) gather of the weights, with cp.async
//                         (16 B, or 4/8 B units for channel granularity 2/4)
//                         written straight into the UMMA swizzle layout;
//                         completion is signalled on mbarriers
//                         (cp.async.mbarrier.arrive), stages are released by
//                         tcgen05.commit.
// Restates (does not port) the conv -> mask -> bn -> relu chains of
// imagenet_classification/models/laud_resnet.py:115-144 of the reference.
#include "laud_common.cuh"

namespace laud {
namespace {

constexpr int BM = 128;                  // UMMA M: output pixels per tile
constexpr int BN_MAX = 256;              // UMMA N upper bound: compact output channels per tile
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * 128;        // 128 rows x 64 fp16
constexpr int B_STAGE_BYTES = BN_MAX * 128;    // 256 rows x 64 fp16
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_WARPS = 4, PROD_WARPS = 4;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int PROD_WARP0 = EPI_WARPS + 1;
constexpr int NUM_THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int KIDX_MAX = 1024;
constexpr uint32_t TMEM_COLS = 512;
constexpr long long SPIN_CYCLES = 4000000000ll;  // watchdog (~2 s): trap instead of hanging the GPU

struct Tables {
  int brow[BN_MAX];        // element offset of the weight row of compact column j (or -1: zero row)
  int kidx[KIDX_MAX];      // this sample's active input-channel groups
  float scale[BN_MAX];
  float shift[BN_MAX];
  int ochan[BN_MAX];       // real output channel of compact column j (or -1: zero pad)
  unsigned long long full[STAGES], empty[STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
};

constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + sizeof(Tables);

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  uint32_t ok = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > SPIN_CYCLES) __trap();
  }
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
// arrive on the mbarrier once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive(unsigned long long* b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate, issued by one thread for the CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(unsigned long long* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;              // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_f16(int n) {
  return (1u << 4)                      // D = fp32
         | (0u << 7) | (0u << 10)       // A, B = fp16
         | (0u << 15) | (0u << 16)      // A, B K-major
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// byte offset of (row r, 16-byte chunk c) inside a swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// ------------------------------------------------------------------ work items
struct Item {
  int b, mt, n0, n_valid, umma_n, Nc, Kc, nk16, cpt, nchunks;
};

__device__ __forceinline__ bool decode_item(const ConvArgs& a, int t, int n_mtiles, int NT, Item& it) {
  const int nt = t % NT;
  const int r = t / NT;
  it.mt = r % n_mtiles;
  const int slot = r / n_mtiles;
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
  const int Nfill = round_up(it.Nc, a.n_pad_align);
  const int ntiles = (Nfill + BN_MAX - 1) / BN_MAX;
  if (nt >= ntiles) return false;
  const int per = round_up((Nfill + ntiles - 1) / ntiles, 16);
  it.n0 = nt * per;
  if (it.n0 >= Nfill) return false;
  it.n_valid = min(per, Nfill - it.n0);
  it.umma_n = round_up(it.n_valid, 16);
  it.Kc = a.k_idx ? __ldg(a.k_cnt + b) * a.k_gran : a.C_in;
  it.nk16 = (it.Kc + 15) >> 4;
  it.cpt = (it.nk16 + 3) >> 2;
  it.nchunks = it.cpt * a.ksize * a.ksize;
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

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_umma_kernel(const __grid_constant__ ConvArgs a, int n_mtiles,
                                                                   int NT, int total_items) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Tables& T = *reinterpret_cast<Tables*>(smem + (size_t)STAGES * STAGE_BYTES);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HWo = a.H_out * a.W_out;
  const int taps = a.ksize * a.ksize;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&T.full[s], PROD_THREADS);
      mbar_init(&T.empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&T.tfull[i], 1);
      mbar_init(&T.tempty[i], EPI_THREADS);
    }
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
    const int pt = threadIdx.x - PROD_WARP0 * 32;      // 0..127
    const int ac = pt & 7, ar0 = pt >> 3;              // A (and dense-K B): 16-byte chunk, first row
    // K-gather unit geometry (bytes per unit, units per 128-byte row)
    const int ub = a.k_idx ? min(16, a.k_gran * 2) : 16;
    const int upc = 128 / ub;
    const int ku = pt % upc, kr0 = pt / upc, krstep = PROD_THREADS / upc;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
      Item it;
      if (!decode_item(a, t, n_mtiles, NT, it)) continue;
      named_bar_sync(1, PROD_THREADS);                 // previous item's table reads are done
      for (int j = pt; j < it.umma_n; j += PROD_THREADS) {
        const int jj = it.n0 + j;
        int off = -1;
        if (jj < it.Nc) {
          const int o = a.n_idx ? __ldg(a.n_idx + (size_t)it.b * a.n_ld + jj / a.n_gran) * a.n_gran + jj % a.n_gran : jj;
          off = o * taps * a.C_in;
        }
        T.brow[j] = off;
      }
      if (a.k_idx) {
        const int ng = it.Kc / a.k_gran;
        for (int q = pt; q < ng; q += PROD_THREADS) T.kidx[q] = __ldg(a.k_idx + (size_t)it.b * a.k_ld + q);
      }
      // the 8 activation rows this thread stages: pixel base and top-left input coordinate
      int pbase[8], iy0[8], ix0[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const RowPos p = row_pos(a, it, ar0 + 16 * i, HWo);
        pbase[i] = p.b * a.H_in * a.W_in;
        iy0[i] = p.valid ? p.oy * a.stride - a.pad : -(1 << 20);
        ix0[i] = p.ox * a.stride - a.pad;
      }
      named_bar_sync(1, PROD_THREADS);
      const int ka_lim = a.k_idx ? it.nk16 * 16 : a.C_in;    // compact inputs are zero-padded to 16 by their producer

      for (int ch = 0; ch < it.nchunks; ++ch) {
        const int tap = ch / it.cpt, kq = ch - tap * it.cpt;
        const int ty = tap / a.ksize, tx = tap - ty * a.ksize;
        const int k0 = kq * 64;
        const int n16 = min(4, it.nk16 - kq * 4);
        const int nch = 2 * n16;                              // 16-byte chunks the MMAs of this stage will read
        mbar_wait(&T.empty[stage], phase ^ 1);
        const uint32_t As = smem_base + stage * STAGE_BYTES;
        const uint32_t Bs = As + A_STAGE_BYTES;
        // ---- A: im2col gather of 128 pixel rows x 64 channels of this tap
        if (ac < nch) {
          const int k = k0 + ac * 8;
          const bool kok = k < ka_lim;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int iy = iy0[i] + ty, ix = ix0[i] + tx;
            const bool ok = kok && iy >= 0 && iy < a.H_in && ix >= 0 && ix < a.W_in;
            const __half* src = ok ? a.x + ((size_t)(pbase[i] + iy * a.W_in + ix)) * a.ldx + k : a.x;
            cp_async_16(As + sw128_off(ar0 + 16 * i, ac), src, ok ? 16u : 0u);
          }
        }
        // ---- B: weight rows of the active output channels, active input channels only
        const __half* wt = a.w + (size_t)tap * a.C_in;
        if (!a.k_idx) {
          if (ac < nch) {
            const int k = k0 + ac * 8;
            const bool kok = k < a.C_in;
            for (int j = ar0; j < it.umma_n; j += 16) {
              const int off = T.brow[j];
              const bool ok = kok && off >= 0;
              cp_async_16(Bs + sw128_off(j, ac), ok ? wt + off + k : a.w, ok ? 16u : 0u);
            }
          }
        } else if (ku * ub < nch * 16) {
          const int e = k0 + ku * (ub >> 1);                   // compact input channel of this unit
          const bool kok = e < it.Kc;
          int col = 0;
          if (kok) col = T.kidx[e / a.k_gran] * a.k_gran + e % a.k_gran;
          const int c16 = (ku * ub) >> 4, w16 = (ku * ub) & 15;
          for (int j = kr0; j < it.umma_n; j += krstep) {
            const int off = T.brow[j];
            const bool ok = kok && off >= 0;
            const uint32_t dst = Bs + sw128_off(j, c16) + w16;
            const __half* src = ok ? wt + off + col : a.w;
            if (ub == 4) cp_async_4(dst, src, ok ? 4u : 0u);
            else if (ub == 8) cp_async_8(dst, src, ok ? 8u : 0u);
            else cp_async_16(dst, src, ok ? 16u : 0u);
          }
        }
        cp_async_arrive(&T.full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == MMA_WARP) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
        Item it;
        if (!decode_item(a, t, n_mtiles, NT, it)) continue;
        mbar_wait(&T.tempty[acc], aphase ^ 1);              // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN_MAX;
        const uint32_t idesc = umma_idesc_f16(it.umma_n);
        for (int ch = 0; ch < it.nchunks; ++ch) {
          const int kq = ch % it.cpt;
          const int n16 = min(4, it.nk16 - kq * 4);
          mbar_wait(&T.full[stage], phase);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t As = smem_base + stage * STAGE_BYTES;
          const uint64_t ad = umma_desc_sw128(As), bd = umma_desc_sw128(As + A_STAGE_BYTES);
          for (int k = 0; k < n16; ++k)                        // +32 B per K=16 step inside the swizzle atom
            umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (ch | k) ? 1u : 0u);
          umma_commit(&T.empty[stage]);                         // frees the stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
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
    int acc = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
      Item it;
      if (!decode_item(a, t, n_mtiles, NT, it)) continue;
      named_bar_sync(2, EPI_THREADS);                           // previous item's table reads are done
      for (int j = et; j < it.umma_n; j += EPI_THREADS) {
        const int jj = it.n0 + j;
        int o = -1;
        float sc = 1.f, sh = 0.f;
        if (jj < it.Nc) {
          o = a.n_idx ? __ldg(a.n_idx + (size_t)it.b * a.n_ld + jj / a.n_gran) * a.n_gran + jj % a.n_gran : jj;
          if (a.scale) { sc = __ldg(a.scale + o); sh = __ldg(a.shift + o); }
        }
        T.ochan[j] = o; T.scale[j] = sc; T.shift[j] = sh;
      }
      const RowPos p = row_pos(a, it, et, HWo);
      const size_t pix = (size_t)p.b * HWo + (size_t)p.oy * a.W_out + p.ox;
      int cls = 0;
      if (a.pre_bias && a.pre_bias_classes > 1)
        cls = border_class(p.oy, a.stride, a.pad, a.H_in) * 4 + border_class(p.ox, a.stride, a.pad, a.W_in);
      const float* pb = a.pre_bias ? a.pre_bias + ((size_t)p.b * a.pre_bias_classes + cls) * a.pre_bias_ld + it.n0 : nullptr;
      const int cpg = a.out_mask ? a.C_out / a.mask_groups : 1;
      named_bar_sync(2, EPI_THREADS);
      mbar_wait(&T.tfull[acc], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN_MAX + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < it.n_valid; c0 += 16) {
        float v[16];
        if (it.nchunks > 0) {
          tmem_ld16(taddr + c0, v);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0.f;
        }
        if (p.valid) {
          if (pb) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 f = __ldg(reinterpret_cast<const float4*>(pb + c0) + q);
              v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w;
            }
          }
          __align__(16) __half rs[16];
          if (a.residual) {
            const uint4* rp = reinterpret_cast<const uint4*>(a.residual + pix * a.ldr + it.n0 + c0);
            *reinterpret_cast<uint4*>(rs) = __ldg(rp);
            if (c0 + 8 < it.n_valid) *reinterpret_cast<uint4*>(rs + 8) = __ldg(rp + 1);
          }
          __align__(16) __half out[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int o = T.ochan[c0 + e];
            float val = v[e] * T.scale[c0 + e] + T.shift[c0 + e];
            int gate = 1;
            if (a.out_mask && o >= 0) {
              gate = a.out_mask[((size_t)p.b * a.mask_groups + o / cpg) * HWo + (size_t)p.oy * a.W_out + p.ox];
              if (a.relu_mode != LAUD_RELU_WHERE_GATE0) val = gate ? val : 0.f;
            }
            if (a.residual && c0 + e < it.n_valid) val += __half2float(rs[e]);
            if (a.relu_mode == LAUD_RELU_ALL || (a.relu_mode == LAUD_RELU_WHERE_GATE0 && !gate)) val = fmaxf(val, 0.f);
            out[e] = __float2half(o >= 0 ? val : 0.f);
          }
          __half* yp = a.y + pix * a.ldy + it.n0 + c0;
          *reinterpret_cast<uint4*>(yp) = *reinterpret_cast<const uint4*>(out);
          if (c0 + 8 < it.n_valid) *reinterpret_cast<uint4*>(yp + 8) = *reinterpret_cast<const uint4*>(out + 8);
        }
      }
      tc_fence_before();
      mbar_arrive(&T.tempty[acc]);
      if (++acc == 2) { acc = 0; aphase ^= 1; }
    }
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
    if (a.k_ld > KIDX_MAX) return false;
    if (a.ldx < round_up(a.C_in, 16)) return false;
  }
  if (a.n_idx && a.n_pad_align < 8) return false;
  if ((size_t)a.C_out * a.ksize * a.ksize * a.C_in >= (1ull << 31)) return false;
  return true;
}

int conv_forward_umma(const ConvArgs& a, cudaStream_t s) {
  if (!conv_umma_supported(a)) return conv_forward_hmma(a, s);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    LAUD_CUDA(cudaGetDevice(&dev));
    LAUD_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    LAUD_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  }
  const long long HWo = (long long)a.H_out * a.W_out;
  const long long rows = a.row_idx ? (long long)a.B * HWo : HWo;
  const int n_mtiles = (int)((rows + BM - 1) / BM);
  const int NT = (round_up(a.C_out, a.n_pad_align) + BN_MAX - 1) / BN_MAX;
  const long long total = (long long)(a.row_idx ? 1 : a.B) * n_mtiles * NT;
  if (total >= (1ll << 31)) {
    set_error("conv_forward_umma: too many tiles");
    return LAUD_E_BADARG;
  }
  const int grid = (int)(total < num_sms ? total : num_sms);
  g_conv_paths[0].fetch_add(1, std::memory_order_relaxed);
  conv_umma_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(a, n_mtiles, NT, (int)total);
  return check_launch("conv_umma_kernel");
}

}  // namespace laud
