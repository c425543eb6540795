// Network ends and bookkeeping kernels: fused stem (conv7x7/2+BN+ReLU+maxpool),
// head (avgpool+fc), H1 constant tables for channel skipping, layout helpers and
// the forward-statistics kernel.
#include "laud_common.cuh"

namespace laud {

// ---------------------------------------------------------------------------
// Stem.  One CTA = one sample x a 4x8 tile of POOLED outputs.  It needs the
// 9x17 conv outputs around it, which need a 23x39 input patch (x3 channels).
// Restates laud_resnet.py:317-324 (conv1 -> bn1 -> relu -> maxpool).
// ---------------------------------------------------------------------------
constexpr int ST_PH = 4, ST_PW = 8;                 // pooled tile
constexpr int ST_CH = 2 * ST_PH + 1, ST_CW = 2 * ST_PW + 1;   // conv tile 9x17
constexpr int ST_IH = 2 * ST_CH + 5, ST_IW = 2 * ST_CW + 5;   // input tile 23x39
constexpr int ST_IWP = ST_IW + 1;

__global__ void __launch_bounds__(256) stem_kernel(const __half* __restrict__ x, int H, int W,
                                                   const __half* __restrict__ w, int C0,
                                                   const float* __restrict__ scale,
                                                   const float* __restrict__ shift,
                                                   __half* __restrict__ y) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_in = reinterpret_cast<float*>(smem_raw);                         // [3][ST_IH][ST_IWP]
  __half* s_w = reinterpret_cast<__half*>(s_in + 3 * ST_IH * ST_IWP);       // [147][C0]
  __half* s_c = s_w + 147 * C0;                                             // [ST_CH*ST_CW][C0]
  const int Hc = H / 2, Wc = W / 2, Hp = Hc / 2, Wp = Wc / 2;
  const int b = blockIdx.z;
  const int py0 = blockIdx.y * ST_PH, px0 = blockIdx.x * ST_PW;
  const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;       // conv-tile origin (may be -1)
  const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;       // input-tile origin
  const int tid = threadIdx.x;

  for (int i = tid; i < 3 * ST_IH * ST_IW; i += 256) {
    const int c = i / (ST_IH * ST_IW), r = (i / ST_IW) % ST_IH, q = i % ST_IW;
    const int iy = iy0 + r, ix = ix0 + q;
    float v = 0.f;
    if (iy >= 0 && ix >= 0 && iy < H && ix < W) v = __half2float(x[(((size_t)b * 3 + c) * H + iy) * W + ix]);
    s_in[(c * ST_IH + r) * ST_IWP + q] = v;
  }
  for (int i = tid; i < 147 * C0; i += 256) {
    const int o = i / 147, t = i % 147;               // global layout [C0][3][7][7]
    s_w[t * C0 + o] = w[i];
  }
  __syncthreads();

  const int ncg = C0 / 8;
  const int items = ST_CH * ST_CW * ncg;
  for (int it = tid; it < items; it += 256) {
    const int cg = it / (ST_CH * ST_CW), pxl = it % (ST_CH * ST_CW);
    const int r = pxl / ST_CW, q = pxl % ST_CW;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int c = 0; c < 3; ++c)
      for (int ky = 0; ky < 7; ++ky) {
        const float* irow = s_in + (c * ST_IH + 2 * r + ky) * ST_IWP + 2 * q;
        const __half* wrow = s_w + ((c * 7 + ky) * 7) * C0 + cg * 8;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float v = irow[kx];
          const uint4 wq = *reinterpret_cast<const uint4*>(wrow + kx * C0);
          const __half2* wh = reinterpret_cast<const __half2*>(&wq);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(wh[e]);
            acc[2 * e] = fmaf(v, f.x, acc[2 * e]);
            acc[2 * e + 1] = fmaf(v, f.y, acc[2 * e + 1]);
          }
        }
      }
    const int cy = cy0 + r, cx = cx0 + q;
    const bool inside = cy >= 0 && cx >= 0 && cy < Hc && cx < Wc;
    __align__(16) __half o8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = cg * 8 + e;
      // outside the conv output: 0 is neutral for the max of post-ReLU values
      const float v = inside ? fmaxf(acc[e] * scale[ch] + shift[ch], 0.f) : 0.f;
      o8[e] = __float2half(v);
    }
    *reinterpret_cast<uint4*>(s_c + (size_t)pxl * C0 + cg * 8) = *reinterpret_cast<uint4*>(o8);
  }
  __syncthreads();

  for (int it = tid; it < ST_PH * ST_PW * ncg; it += 256) {
    const int cg = it % ncg, pp = it / ncg;
    const int pr = pp / ST_PW, pq = pp % ST_PW;
    const int py = py0 + pr, px = px0 + pq;
    if (py >= Hp || px >= Wp) continue;
    __half2 m[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) m[e] = __float2half2_rn(0.f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint4 q4 = *reinterpret_cast<const uint4*>(s_c + (size_t)((2 * pr + dy) * ST_CW + 2 * pq + dx) * C0 + cg * 8);
        const __half2* h = reinterpret_cast<const __half2*>(&q4);
#pragma unroll
        for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], h[e]);
      }
    *reinterpret_cast<uint4*>(y + (((size_t)b * Hp + py) * Wp + px) * C0 + cg * 8) = *reinterpret_cast<uint4*>(m);
  }
}

// ---------------------------------------------------------------------------
// Stem, tensor-core version (the product path for C0 in {16, 32, 64}).
// Persistent CTAs; one work item = one sample x an 8x14 tile of POOLED outputs,
// i.e. a 17x29 tile of conv outputs computed as an implicit GEMM with
// mma.sync.m16n8k16 (fp16 in, fp32 accumulate) straight from the fp16 input
// patch in shared memory:  M = conv pixel, N = output channel,
// K = (c, ky) x 8 with kx fastest (the 8th kx and the 22nd (c,ky) row carry
// zero weights), so an A fragment is a pair of adjacent input pixels - no
// im2col buffer.  BN + ReLU, then the 3x3/2 max-pool from the fp16 conv tile.
// (K = 147 with 3 input channels does not map onto UMMA shared-memory
// descriptors without materialising im2col; at 1.5 % of the network's MACs the
// legacy warp-level MMA is the right tool here.)
// ---------------------------------------------------------------------------
constexpr int S2_PH = 8, S2_PW = 14;
constexpr int S2_CH = 2 * S2_PH + 1, S2_CW = 2 * S2_PW + 1;    // 17 x 29 conv outputs
constexpr int S2_NPIX = S2_CH * S2_CW;                         // 493
constexpr int S2_MT = (S2_NPIX + 15) / 16;                     // 31 m16 tiles
constexpr int S2_IH = 2 * S2_CH + 5, S2_IW = 2 * S2_CW + 5;    // 39 x 63 input patch
constexpr int S2_IWP = 64;                                     // even pitch: fragment loads are 32-bit aligned
constexpr int S2_KS = 11, S2_KP = 184;                         // k16 steps (176 >= 21*8), padded weight-row pitch

__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NT>   // C0 = 8 * NT output channels
__global__ void __launch_bounds__(256, 2) stem_mma_kernel(const __half* __restrict__ x, int B, int H, int W,
                                                          const __half* __restrict__ w,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift,
                                                          __half* __restrict__ y) {
  constexpr int C0 = 8 * NT, CP = C0 + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* s_w = reinterpret_cast<__half*>(smem_raw);            // [C0][S2_KP]
  __half* s_in = s_w + C0 * S2_KP;                              // [3][S2_IH][S2_IWP]
  __half* s_c = s_in + 3 * S2_IH * S2_IWP;                      // [S2_NPIX][CP]
  const int Hc = H / 2, Wc = W / 2, Hp = Hc / 2, Wp = Wc / 2;
  const int tiles_x = (Wp + S2_PW - 1) / S2_PW, tiles_y = (Hp + S2_PH - 1) / S2_PH;
  const int total = B * tiles_x * tiles_y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  __shared__ float2 s_bn[2][C0 / 2];                         // folded BN scale / shift as channel pairs
  for (int i = tid; i < C0; i += 256) {
    reinterpret_cast<float*>(s_bn[0])[i] = scale[i];
    reinterpret_cast<float*>(s_bn[1])[i] = shift[i];
  }
  // weights [C0][3][7][7] -> [C0][(c,ky) x 8 + kx], zero padded
  for (int i = tid; i < C0 * S2_KP; i += 256) {
    const int o = i / S2_KP, k = i - o * S2_KP;
    const int row = k >> 3, kx = k & 7;
    s_w[i] = (row < 21 && kx < 7) ? w[o * 147 + row * 7 + kx] : __float2half(0.f);
  }

  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int b = item / (tiles_x * tiles_y);
    const int tr = item - b * tiles_x * tiles_y;
    const int py0 = (tr / tiles_x) * S2_PH, px0 = (tr % tiles_x) * S2_PW;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;       // conv-tile origin (may be -1)
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;       // input-patch origin
    __syncthreads();                                      // previous item's pooling has finished with s_c / s_in
    if ((W & 1) == 0) {
      // one patch row per warp and pass: ix0 is odd, so patch columns 1.. are pairs of an even-aligned input column and
      // its successor: 32-bit global loads (both inside or both outside the image, W even), 16-bit shared stores
      // all of this warp's rows are requested before the first one is stored: one memory latency per item, not fifteen
      constexpr int RPW = (3 * S2_IH + 7) / 8;               // rows per warp (15)
      __half2 pv[RPW];
      __half p0[RPW];
      const int q = 1 + 2 * lane, ix = ix0 + q;               // lanes 0..30: columns 1..62; lane 31: column 0 (and 63: pad)
      const bool xok = ix >= 0 && ix + 1 < W, x0ok = ix0 >= 0 && ix0 < W;
#pragma unroll
      for (int u = 0; u < RPW; ++u) {
        const int row = warp + 8 * u;
        pv[u] = __float2half2_rn(0.f);
        p0[u] = __float2half(0.f);
        if (row < 3 * S2_IH) {
          const int c = row / S2_IH, rr = row - c * S2_IH;
          const int iy = iy0 + rr;
          if (iy >= 0 && iy < H) {
            const __half* src = x + (((size_t)b * 3 + c) * H + iy) * W;
            if (lane < 31) {
              if (xok) pv[u] = *reinterpret_cast<const __half2*>(src + ix);
            } else if (x0ok) {
              p0[u] = src[ix0];
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < RPW; ++u) {
        const int row = warp + 8 * u;
        if (row < 3 * S2_IH) {
          __half* dst = s_in + row * S2_IWP;
          if (lane < 31) {
            dst[q] = __low2half(pv[u]);
            dst[q + 1] = __high2half(pv[u]);
          } else {
            dst[0] = p0[u];
            dst[S2_IWP - 1] = __float2half(0.f);
          }
        }
      }
    } else {
    for (int i = tid; i < 3 * S2_IH * S2_IWP; i += 256) {
      const int c = i / (S2_IH * S2_IWP), rr = (i / S2_IWP) % S2_IH, q = i % S2_IWP;
      const int iy = iy0 + rr, ix = ix0 + q;
      __half v = __float2half(0.f);
      if (iy >= 0 && ix >= 0 && iy < H && ix < W) v = x[(((size_t)b * 3 + c) * H + iy) * W + ix];
      s_in[i] = v;
    }
    }
    __syncthreads();

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      int poff[2][2];                                     // patch offsets of this thread's 2 pixels in 2 m-tiles
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int p = min((warp + 8 * (2 * pass + m)) * 16 + g + 8 * h, S2_NPIX - 1);
          const int r = p / S2_CW, q = p - r * S2_CW;
          poff[m][h] = 2 * r * S2_IWP + 2 * q + 2 * t;
        }
      float acc[2][NT][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < S2_KS; ++ks) {
        const int ra = 2 * ks, rb = 2 * ks + 1;
        const int offa = ((ra / 7) * S2_IH + ra % 7) * S2_IWP;
        const int offb = rb < 21 ? ((rb / 7) * S2_IH + rb % 7) * S2_IWP : 0;     // pad row: zero weights, any finite data
        uint32_t a[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          a[m][0] = *reinterpret_cast<const uint32_t*>(s_in + poff[m][0] + offa);
          a[m][1] = *reinterpret_cast<const uint32_t*>(s_in + poff[m][1] + offa);
          a[m][2] = *reinterpret_cast<const uint32_t*>(s_in + poff[m][0] + offb);
          a[m][3] = *reinterpret_cast<const uint32_t*>(s_in + poff[m][1] + offb);
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const __half* wp = s_w + (j * 8 + g) * S2_KP + ks * 16 + 2 * t;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wp);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wp + 8);
          mma_16816(acc[0][j], a[0], b0, b1);
          mma_16816(acc[1][j], a[1], b0, b1);
        }
      }
      // BN + ReLU -> fp16 conv tile (0 outside the conv output: neutral for the max of post-ReLU values)
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int p = (warp + 8 * (2 * pass + m)) * 16 + g + 8 * h;
          if (p < S2_NPIX) {
            const int r = p / S2_CW, q = p - r * S2_CW;
            const int cy = cy0 + r, cx = cx0 + q;
            const bool inside = cy >= 0 && cx >= 0 && cy < Hc && cx < Wc;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              const int n = j * 8 + 2 * t;
              float v0 = 0.f, v1 = 0.f;
              if (inside) {
                const float2 sc = s_bn[0][n >> 1], sh = s_bn[1][n >> 1];
                v0 = fmaxf(fmaf(acc[m][j][2 * h], sc.x, sh.x), 0.f);
                v1 = fmaxf(fmaf(acc[m][j][2 * h + 1], sc.y, sh.y), 0.f);
              }
              *reinterpret_cast<__half2*>(s_c + p * CP + n) = __floats2half2_rn(v0, v1);
            }
          }
        }
    }
    __syncthreads();

    constexpr int ncg = C0 / 8;
    for (int it = tid; it < S2_PH * S2_PW * ncg; it += 256) {
      const int cg = it % ncg, pp = it / ncg;
      const int pr = pp / S2_PW, pq = pp - pr * S2_PW;
      const int py = py0 + pr, px = px0 + pq;
      if (py >= Hp || px >= Wp) continue;
      __half2 mx[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) mx[e] = __float2half2_rn(0.f);
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const uint4 q4 = *reinterpret_cast<const uint4*>(s_c + ((2 * pr + dy) * S2_CW + 2 * pq + dx) * CP + cg * 8);
          const __half2* hh = reinterpret_cast<const __half2*>(&q4);
#pragma unroll
          for (int e = 0; e < 4; ++e) mx[e] = __hmax2(mx[e], hh[e]);
        }
      *reinterpret_cast<uint4*>(y + (((size_t)b * Hp + py) * Wp + px) * C0 + cg * 8) = *reinterpret_cast<uint4*>(mx);
    }
  }
}

template <int NT>
int launch_stem_mma(const __half* x, int B, int H, int W, const __half* w, const float* scale, const float* shift,
                    __half* y, cudaStream_t s) {
  constexpr int C0 = 8 * NT;
  const size_t smem = sizeof(__half) * ((size_t)C0 * S2_KP + 3 * S2_IH * S2_IWP + (size_t)S2_NPIX * (C0 + 8));
  static int num_sms_dev[MAX_DEVICES] = {0};
  const int dev = current_device();
  if (num_sms_dev[dev] == 0) {
    int n = 0;
    LAUD_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    LAUD_CUDA(cudaFuncSetAttribute(stem_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    num_sms_dev[dev] = n;
  }
  const int num_sms = num_sms_dev[dev];
  const int Hp = H / 4, Wp = W / 4;
  const long long total = (long long)B * ((Wp + S2_PW - 1) / S2_PW) * ((Hp + S2_PH - 1) / S2_PH);
  const int grid = (int)(total < 2ll * num_sms ? total : 2ll * num_sms);
  stem_mma_kernel<NT><<<grid, 256, smem, s>>>(x, B, H, W, w, scale, shift, y);
  return check_launch("stem_mma_kernel");
}

// ---------------------------------------------------------------------------
// Head fc: CTA = 4 samples x 64 classes; pooled features of the 4 samples in smem.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fc_kernel(const float* __restrict__ pooled, int B, int C,
                                                      const __half* __restrict__ w,
                                                      const float* __restrict__ bias, int n_cls,
                                                      float* __restrict__ logits) {
  extern __shared__ float sp[];   // [4][C]
  const int b0 = blockIdx.y * 4;
  for (int i = threadIdx.x; i < 4 * C; i += 256) {
    const int bb = b0 + i / C;
    sp[i] = bb < B ? pooled[(size_t)bb * C + i % C] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 8; ++k) {
    const int o = blockIdx.x * 64 + warp * 8 + k;
    if (o >= n_cls) break;
    const __half* wr = w + (size_t)o * C;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane * 8; c < C; c += 256) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(wr + c));
      const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          acc[s] = fmaf(f.x, sp[s * C + c + 2 * e], acc[s]);
          acc[s] = fmaf(f.y, sp[s * C + c + 2 * e + 1], acc[s]);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) acc[s] = warp_sum(acc[s]);
    if (lane == 0)
      for (int s = 0; s < 4; ++s)
        if (b0 + s < B) logits[(size_t)(b0 + s) * n_cls + o] = acc[s] + bias[o];
  }
}

// Head fc, tiled: CTA = 32 samples x 64 classes, 128 threads with a 4x4 register tile each; both
// operands staged through shared memory in K-chunks of 32 (fp32 features, fp16 weights widened once).
// The first kernel above re-read the weights once per 4 samples (64x) and paid one LDS per FMA.
constexpr int HF_BS = 32, HF_BO = 64, HF_BK = 32;
__global__ void __launch_bounds__(128) head_fc_tiled_kernel(const float* __restrict__ pooled, int B, int C,
                                                            const __half* __restrict__ w,
                                                            const float* __restrict__ bias, int n_cls,
                                                            float* __restrict__ logits) {
  __shared__ float sa[HF_BK][HF_BS + 4];      // [k][sample]
  __shared__ float sb[HF_BK][HF_BO + 4];      // [k][class]
  const int tid = threadIdx.x;
  const int s0 = blockIdx.y * HF_BS, o0 = blockIdx.x * HF_BO;
  const int ts = (tid & 7) * 4, to = (tid >> 3) * 4;          // this thread's 4 samples x 4 classes
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < C; k0 += HF_BK) {
    for (int i = tid; i < HF_BS * HF_BK; i += 128) {
      const int sidx = i / HF_BK, k = i % HF_BK;               // consecutive threads -> consecutive k: coalesced
      sa[k][sidx] = (s0 + sidx < B && k0 + k < C) ? pooled[(size_t)(s0 + sidx) * C + k0 + k] : 0.f;
    }
    for (int i = tid; i < HF_BO * HF_BK; i += 128) {
      const int o = i / HF_BK, k = i % HF_BK;
      sb[k][o] = (o0 + o < n_cls && k0 + k < C) ? __half2float(w[(size_t)(o0 + o) * C + k0 + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < HF_BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&sa[k][ts]);
      const float4 bv = *reinterpret_cast<const float4*>(&sb[k][to]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (s0 + ts + i < B && o0 + to + j < n_cls)
        logits[(size_t)(s0 + ts + i) * n_cls + o0 + to + j] = acc[i][j] + bias[o0 + to + j];
}

// Head fc on the warp-level tensor cores with fp32-accurate products: the fp32 feature p is split into
// hi = fp16(p) and lo = fp16(p - hi) (p = hi + lo to ~2^-22), the fp16 weight is exact, so
// w*hi + w*lo accumulated in fp32 equals the fp32 product to rounding.  Fragments come straight from
// global memory (features 2 MB, weights 4 MB: L2 resident); warp tile = 16 samples x 32 classes,
// CTA = 4 warps = 64 samples x 32 classes.
__global__ void __launch_bounds__(128) head_fc_mma_kernel(const float* __restrict__ pooled, int B, int C,
                                                          const __half* __restrict__ w,
                                                          const float* __restrict__ bias, int n_cls,
                                                          float* __restrict__ logits) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int s0 = blockIdx.y * 64 + warp * 16, o0 = blockIdx.x * 32;
  if (s0 >= B) return;
  const float* p0 = pooled + (size_t)min(s0 + g, B - 1) * C + 2 * t;          // clamped rows are computed and dropped
  const float* p1 = pooled + (size_t)min(s0 + g + 8, B - 1) * C + 2 * t;
  const __half* wr[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) wr[j] = w + (size_t)min(o0 + j * 8 + g, n_cls - 1) * C + 2 * t;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  // four k16 steps per round: all 48 operand loads of the round are issued before the first MMA consumes one
  auto mma_step = [&](const float2* f, const uint32_t* bq) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 hh = __floats2half2_rn(f[e].x, f[e].y);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(f[e].x - hf.x, f[e].y - hf.y);
      hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mma_16816(acc[j], hi, bq[2 * j], bq[2 * j + 1]);
      mma_16816(acc[j], lo, bq[2 * j], bq[2 * j + 1]);
    }
  };
  int k0 = 0;
  for (; k0 + 64 <= C; k0 += 64) {
    float2 f[4][4];
    uint32_t bq[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = k0 + 16 * u;
      f[u][0] = __ldg(reinterpret_cast<const float2*>(p0 + kk));
      f[u][1] = __ldg(reinterpret_cast<const float2*>(p1 + kk));
      f[u][2] = __ldg(reinterpret_cast<const float2*>(p0 + kk + 8));
      f[u][3] = __ldg(reinterpret_cast<const float2*>(p1 + kk + 8));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bq[u][2 * j] = __ldg(reinterpret_cast<const unsigned int*>(wr[j] + kk));
        bq[u][2 * j + 1] = __ldg(reinterpret_cast<const unsigned int*>(wr[j] + kk + 8));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) mma_step(f[u], bq[u]);
  }
  for (; k0 < C; k0 += 16) {
    float2 f[4];
    uint32_t bq[8];
    f[0] = __ldg(reinterpret_cast<const float2*>(p0 + k0));
    f[1] = __ldg(reinterpret_cast<const float2*>(p1 + k0));
    f[2] = __ldg(reinterpret_cast<const float2*>(p0 + k0 + 8));
    f[3] = __ldg(reinterpret_cast<const float2*>(p1 + k0 + 8));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bq[2 * j] = __ldg(reinterpret_cast<const unsigned int*>(wr[j] + k0));
      bq[2 * j + 1] = __ldg(reinterpret_cast<const unsigned int*>(wr[j] + k0 + 8));
    }
    mma_step(f, bq);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = o0 + j * 8 + 2 * t;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int sidx = s0 + g + (e >> 1) * 8, oo = o + (e & 1);
      if (sidx < B && oo < n_cls) logits[(size_t)sidx * n_cls + oo] = acc[j][e] + bias[oo];
    }
  }
}

// pooled[b, c] = (sum over the tiles of sample b, ascending, of the fused-GAP partial sums) / HW
__global__ void head_pool_partials_kernel(const float* __restrict__ part, int B, int HW, int C, int gap_tiles,
                                          float* __restrict__ pooled) {
  const int b = blockIdx.y;
  const int t_first = (int)(((long long)b * HW) / 128), t_last = (int)(((long long)(b + 1) * HW - 1) / 128);
  const int nk = min(t_last - t_first + 1, gap_tiles);
  const float inv = 1.0f / (float)HW;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int k = 0; k < nk; ++k) v += part[((size_t)b * gap_tiles + k) * C + c];
    pooled[(size_t)b * C + c] = v * inv;
  }
}

// ---------------------------------------------------------------------------
// H1 constants (see laud_b200.h).  The per-sample sums over MASKED channels
//   T2[b,tap,o] = sum_k inact[b,k] * relu(shift1[k]) * w2[o,tap,k]
//   T3[b,o]     = sum_k inact[b,k] * relu(shift2[k]) * w3[o,k]
// are ONE dense GEMM [B x width] x [width x 13*width] on the tcgen05 conv kernel
// (inact = 0/1 indicator of the masked channels, the scaled weights are packed
// at model-preparation time).  What is left here: the 0/1 indicator, and the
// fold of the 9 taps into the 16 border classes of a 3x3/pad-1 convolution.
// ---------------------------------------------------------------------------
__global__ void gate_inactive_kernel(const uint8_t* __restrict__ mask, int B, int G, int gran,
                                     __half* __restrict__ inact) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)B * G * gran;
  if (i >= n) return;
  const int c = (int)(i % (G * gran));
  const long long b = i / (G * gran);
  inact[i] = __float2half(mask[b * G + c / gran] ? 0.f : 1.f);
}

// grid (B), 256 threads.  T: fp16 [B, 9*width + C_out] (row b: T2[tap*width+o], then T3[o]).
__global__ void __launch_bounds__(256) consts_fold_kernel(const __half* __restrict__ T, int width, int C_out,
                                                          const int* __restrict__ idx, const int* __restrict__ cnt,
                                                          int G, int gran, int compact_index,
                                                          float* __restrict__ pre_bias2, float* __restrict__ pre_bias3) {
  const int b = blockIdx.x;
  const __half* Tb = T + (size_t)b * (9 * width + C_out);
  const int n_out = compact_index ? cnt[b] * gran : width;
  for (int j = threadIdx.x; j < n_out; j += 256) {
    const int o = compact_index ? idx[(size_t)b * G + j / gran] * gran + j % gran : j;
    float t[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) t[tap] = __half2float(Tb[tap * width + o]);
#pragma unroll
    for (int cls = 0; cls < 16; ++cls) {
      const int rc = cls >> 2, cc = cls & 3;
      float s = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        if ((dy == 0 && (rc & 1)) || (dy == 2 && (rc & 2))) continue;     // tap row falls outside the input
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          if ((dx == 0 && (cc & 1)) || (dx == 2 && (cc & 2))) continue;
          s += t[dy * 3 + dx];
        }
      }
      pre_bias2[((size_t)b * 16 + cls) * width + j] = s;
    }
  }
  for (int o = threadIdx.x; o < C_out; o += 256) pre_bias3[(size_t)b * C_out + o] = __half2float(Tb[9 * width + o]);
}

// ---------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const void* src, int is_f32, int B, int C, int H, int W, __half* dst, int ldd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)B * C * H * W;
  if (i >= n) return;
  const int c = (int)(i % C);
  const long long p = i / C;             // b*H*W + y*W + x
  const long long hw = (long long)H * W;
  const long long b = p / hw, q = p % hw;
  const long long s = (b * C + c) * hw + q;
  const float v = is_f32 ? reinterpret_cast<const float*>(src)[s] : __half2float(reinterpret_cast<const __half*>(src)[s]);
  dst[p * ldd + c] = __float2half(v);
}
__global__ void nhwc_to_nchw_kernel(const __half* src, int lds, int B, int C, int H, int W, float* dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)B * C * H * W;
  if (i >= n) return;
  const long long hw = (long long)H * W;
  const long long q = i % hw, bc = i / hw;
  const long long b = bc / C, c = bc % C;
  dst[i] = __half2float(src[(b * hw + q) * lds + c]);
}

// ---------------------------------------------------------------------------
// Forward statistics in the reference's evaluation order (single thread).
// consts[i] = {m_chan, m_spat, c1hw, c2hw, c3hw, ds, n_c, n_3, n_2, n_1, flags, extra}
// ---------------------------------------------------------------------------
__global__ void forward_stats_kernel(const int* __restrict__ counts, const long long* __restrict__ consts,
                                     int n_blocks, long long stem_flops, long long pool_flops,
                                     long long fc_flops, float* __restrict__ out) {
  // the warp stages every block's constants and counts in shared memory (loads in flight together); thread 0 then
  // walks the blocks in order - the accumulation itself must stay serial to reproduce the reference's rounding
  constexpr int SMAX = 64;
  __shared__ long long s_k[SMAX * 12];
  __shared__ int s_c[SMAX * 4];
  if (blockIdx.x) return;
  const bool staged = n_blocks <= SMAX;
  if (staged) {
    for (int i = threadIdx.x; i < n_blocks * 12; i += blockDim.x) s_k[i] = consts[i];
    for (int i = threadIdx.x; i < n_blocks * 4; i += blockDim.x) s_c[i] = counts[i];
  }
  __syncthreads();
  if (threadIdx.x) return;
  float flops = 0.f;
  bool flops_is_tensor = false;
  long long flops_int = stem_flops;
  for (int i = 0; i < n_blocks; ++i) {
    const long long* k = staged ? s_k + i * 12 : consts + (size_t)i * 12;
    const int* c = staged ? s_c + i * 4 : counts + (size_t)i * 4;
    const bool use_c = k[10] & 1, use_s = k[10] & 2;
    const float rc = use_c ? __fdiv_rn((float)c[0], (float)k[6]) : 1.0f;
    const float r3 = use_s ? __fdiv_rn((float)c[1], (float)k[7]) : 1.0f;
    const float r2 = use_s ? __fdiv_rn((float)c[2], (float)k[8]) : 1.0f;
    const float r1 = use_s ? __fdiv_rn((float)c[3], (float)k[9]) : 1.0f;
    const long long m = k[0] + k[1];
    float sparse = __fadd_rn((float)m, __fmul_rn(__fmul_rn((float)k[2], rc), r1));
    sparse = __fadd_rn(sparse, __fmul_rn(__fmul_rn((float)k[3], __fmul_rn(rc, rc)), r2));
    sparse = __fadd_rn(sparse, __fmul_rn(__fmul_rn((float)k[4], rc), r3));
    long long dense = m + k[2] + k[3] + k[4];
    const bool regnet_order = (k[10] & 4) != 0;
    if (regnet_order) {
      // laud_regnet.py: flops += se_flops (:195), flops += sparse of the transform (:203), then the block adds the
      // projection to sparse, dense AND flops separately (:285-288)
      if (!flops_is_tensor) flops_int += k[11]; else flops = __fadd_rn(flops, (float)k[11]);
      if (!flops_is_tensor) {
        flops = __fadd_rn((float)flops_int, sparse);
        flops_is_tensor = true;
      } else {
        flops = __fadd_rn(flops, sparse);
      }
      if (k[5]) {
        sparse = __fadd_rn(sparse, (float)k[5]);
        dense += k[5];
        flops = __fadd_rn(flops, (float)k[5]);
      }
    } else {
      if (k[5]) {
        sparse = __fadd_rn(sparse, (float)k[5]);
        dense += k[5];
      }
      if (!flops_is_tensor) {
        flops = __fadd_rn((float)flops_int, sparse);
        flops_is_tensor = true;
      } else {
        flops = __fadd_rn(flops, sparse);
      }
    }
    float* o = out + (size_t)i * 5;
    o[0] = r3; o[1] = r2; o[2] = r1; o[3] = rc;
    o[4] = __fdiv_rn(sparse, (float)dense);
  }
  // avgpool + fc flops are added one at a time as python ints (laud_resnet.py:350,356)
  flops = __fadd_rn(flops, (float)pool_flops);
  flops = __fadd_rn(flops, (float)fc_flops);
  out[(size_t)n_blocks * 5] = flops;
}

}  // namespace laud

using namespace laud;

extern "C" int laud_stem_forward(const void* x, int B, int H, int W, const void* w, int C0, const float* scale,
                                 const float* shift, void* y, void* stream) {
  LAUD_REQUIRE(x && w && scale && shift && y, "laud_stem_forward: null pointer");
  LAUD_REQUIRE(B > 0 && H % 4 == 0 && W % 4 == 0 && C0 % 8 == 0, "laud_stem_forward: need H,W %% 4 == 0, C0 %% 8 == 0");
  {
    const __half *xh = (const __half*)x, *wh = (const __half*)w;
    cudaStream_t st = (cudaStream_t)stream;
    if (C0 == 64) return launch_stem_mma<8>(xh, B, H, W, wh, scale, shift, (__half*)y, st);
    if (C0 == 32) return launch_stem_mma<4>(xh, B, H, W, wh, scale, shift, (__half*)y, st);
    if (C0 == 16) return launch_stem_mma<2>(xh, B, H, W, wh, scale, shift, (__half*)y, st);
  }
  // other stem widths: scalar kernel
  const int Hp = H / 4, Wp = W / 4;
  const size_t smem = sizeof(float) * 3 * ST_IH * ST_IWP + sizeof(__half) * (147 + ST_CH * ST_CW) * (size_t)C0;
  LAUD_REQUIRE(smem <= 200 * 1024, "laud_stem_forward: stem width %d too large", C0);
  static size_t stem_smem_set[MAX_DEVICES] = {0};      // set once per size and device: keeps the launch path free of attribute calls (graph capture)
  const int dev = current_device();
  if (smem > stem_smem_set[dev]) {
    LAUD_CUDA(cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_smem_set[dev] = smem;
  }
  dim3 grid((Wp + ST_PW - 1) / ST_PW, (Hp + ST_PH - 1) / ST_PH, B);
  stem_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const __half*)x, H, W, (const __half*)w, C0, scale, shift,
                                                         (__half*)y);
  return check_launch("stem_kernel");
}

static int launch_head_fc(const float* pooled, int B, int C, const void* w, const float* bias, int n_cls, float* logits,
                          cudaStream_t s) {
  if (C % 16 == 0) {
    head_fc_mma_kernel<<<dim3((n_cls + 31) / 32, (B + 63) / 64), 128, 0, s>>>(pooled, B, C, (const __half*)w, bias, n_cls,
                                                                             logits);
    return check_launch("head_fc_mma_kernel");
  }
  dim3 grid((n_cls + HF_BO - 1) / HF_BO, (B + HF_BS - 1) / HF_BS);
  head_fc_tiled_kernel<<<grid, 128, 0, s>>>(pooled, B, C, (const __half*)w, bias, n_cls, logits);
  return check_launch("head_fc_tiled_kernel");
}

extern "C" int laud_head_forward(const void* x, int B, int HW, int C, const void* w, const float* bias, int n_cls,
                                 float* pooled_ws, float* logits, void* stream) {
  LAUD_REQUIRE(x && w && bias && pooled_ws && logits, "laud_head_forward: null pointer");
  LAUD_REQUIRE(C % 8 == 0 && C <= 8192, "laud_head_forward: need C %% 8 == 0 and C <= 8192 (C=%d)", C);
  // pooled_ws: [B, LAUD_GAP_SPLITS + 1, C]: partial sums followed by the pooled features
  float* pooled = pooled_ws + (size_t)B * LAUD_GAP_SPLITS * C;
  if (int e = laud_global_avg_pool(x, B, HW, C, C, pooled_ws, pooled, stream)) return e;
  return launch_head_fc(pooled, B, C, w, bias, n_cls, logits, (cudaStream_t)stream);
}

extern "C" int laud_head_forward_from_partials(const float* partials, int B, int HW, int C, int gap_tiles, const void* w,
                                               const float* bias, int n_cls, float* pooled_ws, float* logits,
                                               void* stream) {
  LAUD_REQUIRE(partials && w && bias && pooled_ws && logits, "laud_head_forward_from_partials: null pointer");
  LAUD_REQUIRE(B > 0 && HW > 0 && C % 8 == 0 && C <= 8192, "laud_head_forward_from_partials: need C %% 8 == 0 and C <= 8192 (C=%d)", C);
  LAUD_REQUIRE(gap_tiles >= (HW - 1) / 128 + 2, "laud_head_forward_from_partials: gap_tiles %d < (HW-1)/128 + 2", gap_tiles);
  head_pool_partials_kernel<<<dim3((C + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(partials, B, HW, C, gap_tiles,
                                                                                      pooled_ws);
  if (int e = check_launch("head_pool_partials_kernel")) return e;
  return launch_head_fc(pooled_ws, B, C, w, bias, n_cls, logits, (cudaStream_t)stream);
}

extern "C" int laud_gate_inactive(const uint8_t* mask, int B, int G, int gran, void* inact_f16, void* stream) {
  LAUD_REQUIRE(mask && inact_f16 && B > 0 && G > 0 && gran > 0, "laud_gate_inactive: bad arguments");
  const long long n = (long long)B * G * gran;
  gate_inactive_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mask, B, G, gran, (__half*)inact_f16);
  return check_launch("gate_inactive_kernel");
}

extern "C" int laud_channel_consts_fold(const void* T, int B, int width, int C_out, const int32_t* idx,
                                        const int32_t* cnt, int G, int gran, int compact_index, float* pre_bias2,
                                        float* pre_bias3, void* stream) {
  LAUD_REQUIRE(T && idx && cnt && pre_bias2 && pre_bias3, "laud_channel_consts_fold: null pointer");
  LAUD_REQUIRE(G * gran == width, "laud_channel_consts_fold: G*gran (%d*%d) != width %d", G, gran, width);
  consts_fold_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const __half*)T, width, C_out, idx, cnt, G, gran,
                                                         compact_index, pre_bias2, pre_bias3);
  return check_launch("consts_fold_kernel");
}

extern "C" int laud_nchw_to_nhwc_f16(const void* src, int src_is_f32, int B, int C, int H, int W, void* dst, int ldd,
                                     void* stream) {
  LAUD_REQUIRE(src && dst && ldd >= C, "laud_nchw_to_nhwc_f16: bad arguments");
  const long long n = (long long)B * C * H * W;
  nchw_to_nhwc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, src_is_f32, B, C, H, W,
                                                                                    (__half*)dst, ldd);
  return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int laud_nhwc_f16_to_nchw_f32(const void* src, int lds, int B, int C, int H, int W, float* dst,
                                         void* stream) {
  LAUD_REQUIRE(src && dst && lds >= C, "laud_nhwc_f16_to_nchw_f32: bad arguments");
  const long long n = (long long)B * C * H * W;
  nhwc_to_nchw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)src, lds, B, C, H,
                                                                                    W, dst);
  return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int laud_forward_stats(const int32_t* counts, const int64_t* consts, int n_blocks, int64_t stem_flops,
                                  int64_t pool_flops, int64_t fc_flops, float* out, void* stream) {
  LAUD_REQUIRE(counts && consts && out && n_blocks > 0, "laud_forward_stats: bad arguments");
  forward_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counts, (const long long*)consts, n_blocks, stem_flops,
                                                           pool_flops, fc_flops, out);
  return check_launch("forward_stats_kernel");
}
