// Cross-check implementations of the mask-conditioned implicit-GEMM convolution:
//   * conv_forward_naive - one thread per output element, fp32 (self-test oracle on device)
//   * conv_forward_hmma  - warp-level tensor-core (wmma / HMMA) tiles, bring-up path
// The product kernel (tcgen05 + TMEM) lives in conv_umma.cu; all three share
// ConvArgs and the epilogue in laud_common.cuh so they are comparable bit for
// bit in everything but accumulation order.
#include <mma.h>
#include "laud_common.cuh"

namespace laud {

struct RowInfo {
  int b, oy, ox, valid;
};

// Resolve the m-th row of a tile to an output pixel.
__device__ __forceinline__ RowInfo resolve_row(const ConvArgs& a, int slot_b, long long m, int n_rows_s) {
  RowInfo r;
  const int HWo = a.H_out * a.W_out;
  if (a.row_idx) {
    const int cnt = *a.row_cnt;
    r.valid = m < cnt;
    const int flat = r.valid ? a.row_idx[m] : 0;
    r.b = flat / HWo;
    const int p = flat % HWo;
    r.oy = p / a.W_out;
    r.ox = p % a.W_out;
  } else {
    r.valid = m < n_rows_s;
    r.b = slot_b;
    const int p = r.valid ? (int)m : 0;
    r.oy = p / a.W_out;
    r.ox = p % a.W_out;
  }
  return r;
}

__device__ __forceinline__ int real_channel(const int* idx, int ld, int gran, int b, int j) {
  return idx ? idx[(size_t)b * ld + j / gran] * gran + j % gran : j;
}

// --------------------------------------------------------------------------
// naive: grid.x covers rows*Ncols, grid.y = sample slot (or 1 in row mode)
// --------------------------------------------------------------------------
__global__ void conv_naive_kernel(ConvArgs a) {
  const int HWo = a.H_out * a.W_out;
  int b = 0;
  if (!a.row_idx) {
    const int slot = blockIdx.y;
    const int ns = a.sample_cnt ? *a.sample_cnt : a.B;
    if (slot >= ns) return;
    b = a.sample_idx ? a.sample_idx[slot] : slot;
  }
  const int Nmax = round_up(a.C_out, a.n_pad_align);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = t / Nmax;
  const int j = (int)(t % Nmax);
  const RowInfo r = resolve_row(a, b, m, HWo);
  if (!r.valid) return;
  b = r.b;
  const int Kc = a.k_idx ? a.k_cnt[b] * a.k_gran : a.C_in;
  const int Nc = a.n_idx ? a.n_cnt[b] * a.n_gran : a.C_out;
  const size_t pix = ((size_t)b * HWo + (size_t)r.oy * a.W_out + r.ox);
  if (j >= Nc) {
    if (j < round_up(Nc, a.n_pad_align)) a.y[pix * a.ldy + j] = __float2half(0.f);
    return;
  }
  const int o = real_channel(a.n_idx, a.n_ld, a.n_gran, b, j);
  const int taps = a.ksize * a.ksize;
  float acc = 0.f;
  for (int tap = 0; tap < taps; ++tap) {
    const int iy = r.oy * a.stride + tap / a.ksize - a.pad;
    const int ix = r.ox * a.stride + tap % a.ksize - a.pad;
    if (iy < 0 || ix < 0 || iy >= a.H_in || ix >= a.W_in) continue;
    const __half* xp = a.x + (((size_t)b * a.H_in + iy) * a.W_in + ix) * a.ldx;
    const __half* wp = a.w + ((size_t)o * taps + tap) * a.C_in;
    for (int k = 0; k < Kc; ++k) {
      const int kr = real_channel(a.k_idx, a.k_ld, a.k_gran, b, k);
      acc = fmaf(__half2float(xp[k]), __half2float(wp[kr]), acc);
    }
  }
  a.y[pix * a.ldy + j] = __float2half(conv_epilogue(a, acc, b, r.oy, r.ox, j, o));
}

// --------------------------------------------------------------------------
// Low-density fallback for row lists (spatial skipping with almost nothing active): below ~one tensor-core tile of
// active pixels a 128-row MMA tile is mostly padding and the launch is latency-bound; this kernel walks the list with
// one CTA per active pixel: the pixel's input patch (taps x C_in fp16) is staged in shared memory, every thread
// owns output channels and streams their weight rows with 16-byte read-only loads, fp32 accumulate, the shared
// epilogue.  It runs only if row_lo <= *row_cnt < row_hi (device-side dispatch, see ConvArgs).
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_rows_simt_kernel(ConvArgs a) {
  extern __shared__ __align__(16) unsigned char simt_smem[];
  __half* patch = reinterpret_cast<__half*>(simt_smem);          // [taps][C_in]
  const int cnt = __ldg(a.row_cnt);
  if (cnt < a.row_lo || cnt >= a.row_hi) return;
  const int HWo = a.H_out * a.W_out, taps = a.ksize * a.ksize, K = taps * a.C_in;
  for (int m = blockIdx.x; m < cnt; m += gridDim.x) {
    const int flat = __ldg(a.row_idx + m);
    const int b = flat / HWo, p = flat - b * HWo, oy = p / a.W_out, ox = p - oy * a.W_out;
    __syncthreads();                                             // previous pixel's patch no longer read
    for (int i = threadIdx.x * 8; i < K; i += blockDim.x * 8) {
      const int tap = i / a.C_in, k = i - tap * a.C_in;
      const int iy = oy * a.stride + tap / a.ksize - a.pad, ix = ox * a.stride + tap % a.ksize - a.pad;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (iy >= 0 && ix >= 0 && iy < a.H_in && ix < a.W_in)
        v = __ldg(reinterpret_cast<const uint4*>(a.x + (((size_t)b * a.H_in + iy) * a.W_in + ix) * a.ldx + k));
      *reinterpret_cast<uint4*>(patch + i) = v;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < a.C_out; o += blockDim.x) {
      const __half* wrow = a.w + (size_t)o * K;
      float acc = 0.f;
      for (int i = 0; i < K; i += 8) {
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wrow + i));
        const uint4 xv = *reinterpret_cast<const uint4*>(patch + i);
        const __half2* wh = reinterpret_cast<const __half2*>(&wv);
        const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 wf = __half22float2(wh[e]), xf = __half22float2(xh[e]);
          acc = fmaf(xf.x, wf.x, acc);
          acc = fmaf(xf.y, wf.y, acc);
        }
      }
      a.y[((size_t)b * HWo + p) * a.ldy + o] = __float2half(conv_epilogue(a, acc, b, oy, ox, o, o));
    }
  }
}

bool conv_rows_simt_supported(const ConvArgs& a) {
  return a.row_idx && !a.k_idx && !a.n_idx && !a.n_mask && !a.pre_bias && !a.bias_t && !a.sample_idx && !a.gap_partial &&
         a.C_in % 8 == 0 && a.ldx % 8 == 0 && (size_t)a.ksize * a.ksize * a.C_in * 2 <= 96 * 1024;
}

int conv_forward_rows_simt(const ConvArgs& a, cudaStream_t s) {
  const size_t smem = (size_t)a.ksize * a.ksize * a.C_in * sizeof(__half);
  static size_t smem_set[MAX_DEVICES] = {0};
  const int dev = current_device();
  if (smem > 48 * 1024 && smem > smem_set[dev]) {
    LAUD_CUDA(cudaFuncSetAttribute(conv_rows_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[dev] = smem;
  }
  const int grid = a.row_hi < 1024 ? (a.row_hi > 0 ? a.row_hi : 1) : 1024;   // at most row_hi - 1 pixels are ever processed here
  conv_rows_simt_kernel<<<grid, 256, smem, s>>>(a);
  return check_launch("conv_rows_simt_kernel");
}

int conv_forward_naive(const ConvArgs& a, cudaStream_t s) {
  const int Nmax = round_up(a.C_out, a.n_pad_align);
  const long long HWo = (long long)a.H_out * a.W_out;
  const long long rows = a.row_idx ? (long long)a.B * HWo : HWo;
  const long long threads = rows * Nmax;
  dim3 grid((unsigned)((threads + 255) / 256), a.row_idx ? 1 : a.B);
  g_conv_paths[2].fetch_add(1, std::memory_order_relaxed);
  conv_naive_kernel<<<grid, 256, 0, s>>>(a);
  return check_launch("conv_naive_kernel");
}

// --------------------------------------------------------------------------
// HMMA (wmma) tiles: 64x64x32 per CTA, 4 warps of 32x32.
// K is enumerated as (tap, compact channel) with the per-tap extent padded to
// a multiple of 8 (Kp): the producer of a compact activation zero-fills the
// pad channels, the weight gather returns 0 for them.
// --------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 32, LDA = BK + 8, LDB = BK + 8, LDC = BN + 4;

__device__ __forceinline__ uint4 load_w_vec(const ConvArgs& a, int b, int o, int tap, int taps, int j, int Kc) {
  // 8 consecutive compact in-channels [j, j+8) of weight row (o, tap)
  const __half* wrow = a.w + ((size_t)o * taps + tap) * a.C_in;
  if (!a.k_idx) return __ldg(reinterpret_cast<const uint4*>(wrow + j));
  __align__(16) __half tmp[8];
  const int* kl = a.k_idx + (size_t)b * a.k_ld;
  if (a.k_gran % 8 == 0) {
    return j < Kc ? __ldg(reinterpret_cast<const uint4*>(wrow + kl[j / a.k_gran] * a.k_gran + j % a.k_gran))
                  : make_uint4(0, 0, 0, 0);
  } else if (a.k_gran % 2 == 0) {
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const int jj = j + e;
      __half2 v = __float2half2_rn(0.f);
      if (jj < Kc) v = __ldg(reinterpret_cast<const __half2*>(wrow + kl[jj / a.k_gran] * a.k_gran + jj % a.k_gran));
      *reinterpret_cast<__half2*>(&tmp[e]) = v;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int jj = j + e;
      tmp[e] = jj < Kc ? wrow[kl[jj / a.k_gran] * a.k_gran + jj % a.k_gran] : __float2half(0.f);
    }
  }
  return *reinterpret_cast<uint4*>(tmp);
}

__global__ void __launch_bounds__(128) conv_hmma_kernel(ConvArgs a) {
  using namespace nvcuda;
  __shared__ __align__(32) __half As[BM * LDA];
  __shared__ __align__(32) __half Bs[BN * LDB];
  __shared__ __align__(32) float Cs[BM * LDC];

  const int HWo = a.H_out * a.W_out;
  const int tid = threadIdx.x, warp = tid >> 5;
  int b = 0;
  if (!a.row_idx) {
    const int slot = blockIdx.y;
    const int ns = a.sample_cnt ? *a.sample_cnt : a.B;
    if (slot >= ns) return;
    b = a.sample_idx ? a.sample_idx[slot] : slot;
  } else {
    if ((long long)blockIdx.x * BM >= *a.row_cnt) return;
  }
  const int Kc = a.k_idx ? a.k_cnt[b] * a.k_gran : a.C_in;
  const int Kp = (Kc + 7) & ~7;
  const int Nc = a.n_idx ? a.n_cnt[b] * a.n_gran : a.C_out;
  const int Nfill = round_up(Nc, a.n_pad_align);
  const int n0 = blockIdx.z * BN;
  if (n0 >= Nfill) return;
  const int taps = a.ksize * a.ksize;
  const int Ktot = taps * Kp;
  const long long m0 = (long long)blockIdx.x * BM;

  // loader mapping: 2 rows x 1 vector per thread for A and for B
  const int lvec = tid & 3, lrow = tid >> 2;       // rows lrow, lrow+32
  RowInfo ri[2];
  int wo[2];                                       // real out channel of the B rows (or -1)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    ri[i] = resolve_row(a, b, m0 + lrow + 32 * i, HWo);
    const int j = n0 + lrow + 32 * i;
    wo[i] = j < Nc ? real_channel(a.n_idx, a.n_ld, a.n_gran, b, j) : -1;
  }

  wmma::fragment<wmma::accumulator, 16, 16, 16, float> fc[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) wmma::fill_fragment(fc[i][j], 0.f);
  const int wm = warp >> 1, wn = warp & 1;

  for (int k0 = 0; k0 < Ktot; k0 += BK) {
    const int kk = k0 + lvec * 8;
    const int tap = kk / Kp, j = kk - tap * Kp;
    const int ty = tap / a.ksize, tx = tap - ty * a.ksize;
    uint4 va[2], vb[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      va[i] = make_uint4(0, 0, 0, 0);
      vb[i] = make_uint4(0, 0, 0, 0);
      if (kk < Ktot) {
        const int iy = ri[i].oy * a.stride + ty - a.pad, ix = ri[i].ox * a.stride + tx - a.pad;
        if (ri[i].valid && iy >= 0 && ix >= 0 && iy < a.H_in && ix < a.W_in)
          va[i] = __ldg(reinterpret_cast<const uint4*>(
              a.x + (((size_t)ri[i].b * a.H_in + iy) * a.W_in + ix) * a.ldx + j));
        if (wo[i] >= 0) vb[i] = load_w_vec(a, b, wo[i], tap, taps, j, Kc);
      }
    }
    __syncthreads();   // previous iteration's fragments are consumed
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      *reinterpret_cast<uint4*>(&As[(lrow + 32 * i) * LDA + lvec * 8]) = va[i];
      *reinterpret_cast<uint4*>(&Bs[(lrow + 32 * i) * LDB + lvec * 8]) = vb[i];
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < BK; ks += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> fa[2];
      wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> fb[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        wmma::load_matrix_sync(fa[i], &As[(wm * 32 + i * 16) * LDA + ks], LDA);
        wmma::load_matrix_sync(fb[i], &Bs[(wn * 32 + i * 16) * LDB + ks], LDB);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j2 = 0; j2 < 2; ++j2) wmma::mma_sync(fc[i][j2], fa[i], fb[j2], fc[i][j2]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j2 = 0; j2 < 2; ++j2)
      wmma::store_matrix_sync(&Cs[(wm * 32 + i * 16) * LDC + wn * 32 + j2 * 16], fc[i][j2], LDC,
                              wmma::mem_row_major);
  __syncthreads();

  // epilogue: thread -> (row = tid/32 + 4*i, column pair = (tid%32)*2)
  const int cp = (tid & 31) * 2;
  for (int r = tid >> 5; r < BM; r += 4) {
    const RowInfo q = resolve_row(a, b, m0 + r, HWo);
    if (!q.valid) continue;
    const size_t pix = (size_t)q.b * HWo + (size_t)q.oy * a.W_out + q.ox;
    float v[2];
    bool wr[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = n0 + cp + e;
      wr[e] = j < Nfill;
      v[e] = 0.f;
      if (j < Nc) {
        const int o = real_channel(a.n_idx, a.n_ld, a.n_gran, q.b, j);
        v[e] = conv_epilogue(a, Cs[r * LDC + cp + e], q.b, q.oy, q.ox, j, o);
      }
    }
    __half* yp = a.y + pix * a.ldy + n0 + cp;
    if (wr[0] && wr[1]) *reinterpret_cast<__half2*>(yp) = __floats2half2_rn(v[0], v[1]);
    else if (wr[0]) yp[0] = __float2half(v[0]);
  }
}

int conv_forward_hmma(const ConvArgs& a, cudaStream_t s) {
  const long long HWo = (long long)a.H_out * a.W_out;
  const long long rows = a.row_idx ? (long long)a.B * HWo : HWo;
  const int Nmax = round_up(a.C_out, a.n_pad_align);
  dim3 grid((unsigned)((rows + BM - 1) / BM), a.row_idx ? 1 : a.B, (Nmax + BN - 1) / BN);
  g_conv_paths[1].fetch_add(1, std::memory_order_relaxed);
  conv_hmma_kernel<<<grid, 128, 0, s>>>(a);
  return check_launch("conv_hmma_kernel");
}

}  // namespace laud
