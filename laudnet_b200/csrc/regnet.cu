// LAUD-RegNet-Y specific operators (imagenet_classification/models/laud_regnet.py of the reference): the 3x3/2
// stem, the grouped 3x3 convolution of the bottleneck transform (group width 8/16: far too narrow for a tensor-core
// tile, so it is a CUDA-core kernel that keeps one group's weights in shared memory), and Squeeze-Excitation.
// The 1x1 convolutions (a, c, proj), the maskers and the head run on the kernels shared with LAUD-ResNet.
#include <string.h>
#include "laud_common.cuh"

namespace laud {

// ---------------------------------------------------------------------------
// SimpleStemIN: conv3x3/2 (pad 1) + BN + ReLU.  x fp16 NCHW [B,3,H,W] -> y fp16 NHWC [B,H/2,W/2,C0].
// One thread = one output pixel (its 27 inputs in registers, all C0 channels); the 27 x C0 weights sit in shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) regnet_stem_kernel(const __half* __restrict__ x, int B, int H, int W,
                                                          const __half* __restrict__ w, int C0,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift, __half* __restrict__ y) {
  extern __shared__ __align__(16) float s_w[];         // [27][C0]
  for (int i = threadIdx.x; i < 27 * C0; i += 256) {
    const int o = i / 27, t = i % 27;                  // global layout [C0][3][3][3]
    s_w[t * C0 + o] = __half2float(w[i]);
  }
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2, ncg = C0 / 8;
  const long long total = (long long)B * Ho * Wo;       // one thread = one output pixel, all channels
  for (long long pix = (long long)blockIdx.x * 256 + threadIdx.x; pix < total; pix += (long long)gridDim.x * 256) {
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
    float in[27];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
          const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
          in[(c * 3 + ky) * 3 + kx] = ok ? __half2float(x[(((size_t)b * 3 + c) * H + iy) * W + ix]) : 0.f;
        }
    for (int cg = 0; cg < ncg; ++cg) {
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
      for (int t = 0; t < 27; ++t) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_w + t * C0 + cg * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(s_w + t * C0 + cg * 8 + 4);
        acc[0] = fmaf(in[t], w0.x, acc[0]); acc[1] = fmaf(in[t], w0.y, acc[1]);
        acc[2] = fmaf(in[t], w0.z, acc[2]); acc[3] = fmaf(in[t], w0.w, acc[3]);
        acc[4] = fmaf(in[t], w1.x, acc[4]); acc[5] = fmaf(in[t], w1.y, acc[5]);
        acc[6] = fmaf(in[t], w1.z, acc[6]); acc[7] = fmaf(in[t], w1.w, acc[7]);
      }
      __align__(16) __half o8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ch = cg * 8 + e;
        o8[e] = __float2half(fmaxf(fmaf(acc[e], scale[ch], shift[ch]), 0.f));
      }
      *reinterpret_cast<uint4*>(y + (size_t)pix * C0 + cg * 8) = *reinterpret_cast<const uint4*>(o8);
    }
  }
}

// ---------------------------------------------------------------------------
// Grouped 3x3 convolution (pad 1, stride 1|2) + BN + ReLU, NHWC fp16, group width GW (8 | 16).
//   w fp16 [C][9][GW]  (out channel, tap, in channel of the group)
//   in_mask (optional) u8 [B, C/in_gran]: channel gate applied to the INPUT (the reference masks conv a's output
//   after BN+ReLU, laud_regnet.py:183 - i.e. zeros);  out_gate (optional) the same gate applied to the OUTPUT (:189).
// grid (pixel tiles, groups); one thread = one output pixel x 8 output channels of the group.
// ---------------------------------------------------------------------------
template <int GW>
__global__ void __launch_bounds__(128) grouped_conv3x3_kernel(const __half* __restrict__ x, int B, int H_in, int W_in,
                                                              int C, int stride, const __half* __restrict__ w,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const uint8_t* __restrict__ ch_mask, int mask_gran,
                                                              __half* __restrict__ y) {
  __shared__ float s_w[GW * 9 * GW];                   // [out in group][tap][in]
  const int grp = blockIdx.y, c0 = grp * GW;
  for (int i = threadIdx.x; i < GW * 9 * GW; i += 128) s_w[i] = __half2float(w[(size_t)c0 * 9 * GW + i]);
  __syncthreads();
  const int H_out = H_in / stride, W_out = W_in / stride;
  constexpr int HALVES = GW / 8;                        // 8-channel output slices per group
  const long long total = (long long)B * H_out * W_out * HALVES;
  const int G_mask = ch_mask ? C / mask_gran : 0;
  for (long long it = (long long)blockIdx.x * 128 + threadIdx.x; it < total; it += (long long)gridDim.x * 128) {
    const int hf = (int)(it % HALVES);
    const long long pix = it / HALVES;
    const int ox = (int)(pix % W_out), oy = (int)((pix / W_out) % H_out), b = (int)(pix / ((long long)W_out * H_out));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    float gin[GW];                                      // input gate of the group's channels (1 when no mask)
#pragma unroll
    for (int k = 0; k < GW; ++k) gin[k] = ch_mask ? (float)ch_mask[(size_t)b * G_mask + (c0 + k) / mask_gran] : 1.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * stride - 1 + ky;
      if (iy < 0 || iy >= H_in) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * stride - 1 + kx;
        if (ix < 0 || ix >= W_in) continue;
        const __half* xp = x + (((size_t)b * H_in + iy) * W_in + ix) * C + c0;
        float xin[GW];
#pragma unroll
        for (int v = 0; v < GW / 8; ++v) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(xp) + v);
          const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h2[e]);
            xin[v * 8 + 2 * e] = f.x * gin[v * 8 + 2 * e];
            xin[v * 8 + 2 * e + 1] = f.y * gin[v * 8 + 2 * e + 1];
          }
        }
        const int tap = ky * 3 + kx;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float* wr = s_w + ((hf * 8 + e) * 9 + tap) * GW;
#pragma unroll
          for (int k = 0; k < GW; ++k) acc[e] = fmaf(xin[k], wr[k], acc[e]);
        }
      }
    }
    __align__(16) __half o8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = c0 + hf * 8 + e;
      float v = fmaxf(fmaf(acc[e], scale[ch], shift[ch]), 0.f);
      if (ch_mask) v *= (float)ch_mask[(size_t)b * G_mask + ch / mask_gran];
      o8[e] = __float2half(v);
    }
    *reinterpret_cast<uint4*>(y + (size_t)pix * C + c0 + hf * 8) = *reinterpret_cast<const uint4*>(o8);
  }
}

// ---------------------------------------------------------------------------
// Grouped 3x3, group width 16, no channel gate: warp-level tensor-core version (the product path of RegNetY-800MF).
// Per group the convolution is a GEMM  [pixels] x [16 out] x [K = 9 taps x 16 in]: one tap is exactly one k16 step of
// mma.sync.m16n8k16, whose A fragment is 16 pixels x the group's 16 contiguous input channels (32 bytes per pixel,
// read straight from global memory / L1 - no im2col buffer) and whose B fragments come from the group's weights in
// shared memory.  grid (pixel tiles, groups); a warp owns 16 output pixels per iteration.
// (Group width 16 cannot fill a tcgen05 tile: N = 16 at M = 128 would cost the same ~128 cycles as N = 256.)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816_f16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                              uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int GC_WP = 24;      // padded k pitch (halves) of a weight row in shared memory: conflict-free fragment loads
__global__ void __launch_bounds__(128) grouped_conv3x3_mma16_kernel(const __half* __restrict__ x, int B, int H_in,
                                                                    int W_in, int C, int stride,
                                                                    const __half* __restrict__ w,
                                                                    const float* __restrict__ scale,
                                                                    const float* __restrict__ shift,
                                                                    __half* __restrict__ y) {
  __shared__ __align__(16) __half s_w[9 * 16 * GC_WP];           // [tap][out n][k]
  const int grp = blockIdx.y, c0 = grp * 16;
  for (int i = threadIdx.x; i < 16 * 9 * 16; i += 128) {         // global [out][tap][in]
    const int n = i / 144, r = i % 144, tap = r / 16, k = r % 16;
    s_w[(tap * 16 + n) * GC_WP + k] = w[(size_t)c0 * 144 + i];
  }
  __syncthreads();
  const int H_out = H_in / stride, W_out = W_in / stride;
  const long long npix = (long long)B * H_out * W_out;
  const long long ntiles = (npix + 15) / 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  float sc[4], sh[4];                                            // this thread's 4 output channels: nt*8 + 2t, +1
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    sc[2 * nt] = scale[c0 + nt * 8 + 2 * t];
    sc[2 * nt + 1] = scale[c0 + nt * 8 + 2 * t + 1];
    sh[2 * nt] = shift[c0 + nt * 8 + 2 * t];
    sh[2 * nt + 1] = shift[c0 + nt * 8 + 2 * t + 1];
  }
  for (long long tile = (long long)blockIdx.x * 4 + warp; tile < ntiles; tile += (long long)gridDim.x * 4) {
    int oy[2], ox[2], bb[2];
    bool pv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      long long p = tile * 16 + g + 8 * h;
      pv[h] = p < npix;
      if (!pv[h]) p = npix - 1;
      ox[h] = (int)(p % W_out);
      oy[h] = (int)((p / W_out) % H_out);
      bb[h] = (int)(p / ((long long)W_out * H_out));
    }
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      uint32_t a[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int iy = oy[h] * stride - 1 + ky, ix = ox[h] * stride - 1 + kx;
        const bool ok = iy >= 0 && iy < H_in && ix >= 0 && ix < W_in;
        uint32_t lo = 0u, hi = 0u;
        if (ok) {
          const uint32_t* xp = reinterpret_cast<const uint32_t*>(x + (((size_t)bb[h] * H_in + iy) * W_in + ix) * C + c0);
          lo = __ldg(xp + t);                                    // channels 2t, 2t+1
          hi = __ldg(xp + 4 + t);                                // channels 2t+8, 2t+9
        }
        a[h] = lo;                                               // a0 (row g) / a1 (row g+8)
        a[2 + h] = hi;                                           // a2 / a3
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const __half* wp = s_w + (tap * 16 + nt * 8 + g) * GC_WP + 2 * t;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wp);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wp + 8);
        mma_16816_f16(acc[nt], a[0], a[1], a[2], a[3], b0, b1);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!pv[h]) continue;
      const long long p = tile * 16 + g + 8 * h;
      __half* yp = y + (size_t)p * C + c0 + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float v0 = fmaxf(fmaf(acc[nt][2 * h], sc[2 * nt], sh[2 * nt]), 0.f);
        const float v1 = fmaxf(fmaf(acc[nt][2 * h + 1], sc[2 * nt + 1], sh[2 * nt + 1]), 0.f);
        *reinterpret_cast<__half2*>(yp + nt * 8) = __floats2half2_rn(v0, v1);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Squeeze-Excitation gate (torchvision SqueezeExcitation, laud_regnet.py:194): per sample
//   s = sigmoid(W2 relu(W1 p + b1) + b2),  p = pooled features [C], W1 [S][C], W2 passed transposed as [S][C]  (optionally gated: p *= mask, s *= mask - the
//   channel gate of :189 commutes with the pooling because it is constant over the pixels).
// One CTA per sample.  dynamic smem: p[C] | h[S]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ pooled, int C, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, int S, const float* __restrict__ w2,
                                                      const float* __restrict__ b2, const uint8_t* __restrict__ ch_mask,
                                                      int mask_gran, float* __restrict__ gate) {
  extern __shared__ float sm[];
  float* p = sm;
  float* h = sm + C;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G_mask = ch_mask ? C / mask_gran : 0;
  for (int c = tid; c < C; c += 256) {
    float v = pooled[(size_t)b * C + c];
    if (ch_mask) v *= (float)ch_mask[(size_t)b * G_mask + c / mask_gran];
    p[c] = v;
  }
  __syncthreads();
  for (int j = warp; j < S; j += 8) {
    const float* wr = w1 + (size_t)j * C;
    float t = 0.f;
    for (int c = lane; c < C; c += 32) t = fmaf(wr[c], p[c], t);
    t = warp_sum(t);
    if (lane == 0) h[j] = fmaxf(t + b1[j], 0.f);
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {                 // w2 is stored TRANSPOSED [S][C]: coalesced across the threads
    float t = 0.f;
    for (int j = 0; j < S; ++j) t = fmaf(__ldg(w2 + (size_t)j * C + c), h[j], t);
    float s = 1.f / (1.f + __expf(-(t + b2[c])));
    if (ch_mask) s *= (float)ch_mask[(size_t)b * G_mask + c / mask_gran];
    gate[(size_t)b * C + c] = s;
  }
}

// x[b, p, c] *= gate[b, c]  (fp16 NHWC in place, 16 bytes per thread)
__global__ void __launch_bounds__(256) scale_channels_kernel(__half* __restrict__ x, long long n_vec, int HW, int C,
                                                             const float* __restrict__ gate) {
  const int nvc = C / 8;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_vec; i += (long long)gridDim.x * 256) {
    const int cv = (int)(i % nvc);
    const long long b = i / ((long long)nvc * HW);
    uint4 q = *reinterpret_cast<const uint4*>(x + i * 8);
    __half2* h2 = reinterpret_cast<__half2*>(&q);
    const float* g = gate + b * C + cv * 8;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __half22float2(h2[e]);
      f.x *= g[2 * e];
      f.y *= g[2 * e + 1];
      h2[e] = __floats2half2_rn(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(x + i * 8) = q;
  }
}

}  // namespace laud

using namespace laud;

extern "C" int laud_regnet_stem_forward(const void* x_nchw, int B, int H, int W, const void* w, int C0,
                                        const float* scale, const float* shift, void* y_nhwc, void* stream) {
  LAUD_REQUIRE(x_nchw && w && scale && shift && y_nhwc, "laud_regnet_stem_forward: null pointer");
  LAUD_REQUIRE(B > 0 && H % 2 == 0 && W % 2 == 0 && C0 % 8 == 0 && C0 <= 256,
               "laud_regnet_stem_forward: need even H,W and C0 %% 8 == 0, C0 <= 256");
  const long long total = (long long)B * (H / 2) * (W / 2);
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  regnet_stem_kernel<<<grid, 256, sizeof(float) * 27 * C0, (cudaStream_t)stream>>>(
      (const __half*)x_nchw, B, H, W, (const __half*)w, C0, scale, shift, (__half*)y_nhwc);
  return check_launch("regnet_stem_kernel");
}

extern "C" int laud_grouped_conv3x3_forward(const void* x, int B, int H_in, int W_in, int C, int stride, const void* w,
                                            int group_width, const float* scale, const float* shift,
                                            const uint8_t* ch_mask, int mask_gran, void* y, void* stream) {
  LAUD_REQUIRE(x && w && scale && shift && y, "laud_grouped_conv3x3_forward: null pointer");
  LAUD_REQUIRE(B > 0 && (stride == 1 || stride == 2) && H_in % stride == 0 && W_in % stride == 0,
               "laud_grouped_conv3x3_forward: stride must be 1 or 2 and divide H,W");
  LAUD_REQUIRE(group_width >= 8 && group_width % 8 == 0 && C % group_width == 0,
               "laud_grouped_conv3x3_forward: group width must be a multiple of 8 and divide C (got %d, C=%d)", group_width, C);
  LAUD_REQUIRE(!ch_mask || (mask_gran >= 1 && C % mask_gran == 0), "laud_grouped_conv3x3_forward: bad mask granularity");
  if (group_width != 8 && group_width != 16 && group_width != 24) {
    // Wide groups (RegNetY-8GF / 16GF / 32GF: 56, 112, 232 channels per group) fill tensor-core tiles: every group is an
    // ordinary 3x3 convolution over a channel slice of the NHWC tensors (pitch C) with its own [gw, 9, gw] rows of w,
    // run on the mask-conditioned convolution kernel.
    if (ch_mask) {
      set_error("laud_grouped_conv3x3_forward: the channel gate is built for group widths 8 / 16 / 24 only");
      return LAUD_E_UNSUPPORTED;
    }
    const int groups_w = C / group_width;
    for (int g = 0; g < groups_w; ++g) {
      laud_conv_desc d;
      memset(&d, 0, sizeof(d));
      d.x = (const __half*)x + (size_t)g * group_width;  d.ldx = C;
      d.w = (const __half*)w + (size_t)g * group_width * 9 * group_width;
      d.y = (__half*)y + (size_t)g * group_width;        d.ldy = C;
      d.B = B; d.H_in = H_in; d.W_in = W_in; d.C_in = group_width;
      d.H_out = H_in / stride; d.W_out = W_in / stride; d.C_out = group_width;
      d.ksize = 3; d.stride = stride; d.pad = 1;
      d.scale = scale + (size_t)g * group_width; d.shift = shift + (size_t)g * group_width;
      d.relu_mode = LAUD_RELU_ALL;
      d.mask_groups = 1;
      const int rc = laud_conv_forward(&d, LAUD_CONV_AUTO, stream);
      if (rc != LAUD_OK) return rc;
    }
    return LAUD_OK;
  }
  const long long total = (long long)B * (H_in / stride) * (W_in / stride) * (group_width / 8);
  const int groups = C / group_width;
  long long tiles = (total + 127) / 128;
  const long long cap = (148ll * 32 + groups - 1) / groups;        // ~32 CTAs per SM worth of work items overall
  if (tiles > cap) tiles = cap;
  dim3 grid((unsigned)tiles, (unsigned)groups);
  cudaStream_t s = (cudaStream_t)stream;
  if (group_width == 16 && !ch_mask) {
    const long long ntiles = ((long long)B * (H_in / stride) * (W_in / stride) + 15) / 16;
    long long gx = (ntiles + 3) / 4;
    if (gx > cap) gx = cap;
    grouped_conv3x3_mma16_kernel<<<dim3((unsigned)gx, (unsigned)groups), 128, 0, s>>>(
        (const __half*)x, B, H_in, W_in, C, stride, (const __half*)w, scale, shift, (__half*)y);
    return check_launch("grouped_conv3x3_mma16_kernel");
  }
  if (group_width == 24)
    grouped_conv3x3_kernel<24><<<grid, 128, 0, s>>>((const __half*)x, B, H_in, W_in, C, stride, (const __half*)w, scale,
                                                    shift, ch_mask, mask_gran, (__half*)y);
  else if (group_width == 16)
    grouped_conv3x3_kernel<16><<<grid, 128, 0, s>>>((const __half*)x, B, H_in, W_in, C, stride, (const __half*)w, scale,
                                                    shift, ch_mask, mask_gran, (__half*)y);
  else
    grouped_conv3x3_kernel<8><<<grid, 128, 0, s>>>((const __half*)x, B, H_in, W_in, C, stride, (const __half*)w, scale,
                                                   shift, ch_mask, mask_gran, (__half*)y);
  return check_launch("grouped_conv3x3_kernel");
}

extern "C" int laud_se_gate(const float* pooled, int B, int C, const float* w1, const float* b1, int S, const float* w2,
                            const float* b2, const uint8_t* ch_mask, int mask_gran, float* gate, void* stream) {
  LAUD_REQUIRE(pooled && w1 && b1 && w2 && b2 && gate && B > 0 && C > 0 && S > 0, "laud_se_gate: bad arguments");
  const size_t smem = sizeof(float) * ((size_t)C + S);
  LAUD_REQUIRE(smem <= 48 * 1024, "laud_se_gate: C + S too large");
  se_gate_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(pooled, C, w1, b1, S, w2, b2, ch_mask, mask_gran, gate);
  return check_launch("se_gate_kernel");
}

extern "C" int laud_scale_channels(void* x, int B, int HW, int C, const float* gate, void* stream) {
  LAUD_REQUIRE(x && gate && B > 0 && HW > 0 && C % 8 == 0, "laud_scale_channels: need C %% 8 == 0");
  const long long n_vec = (long long)B * HW * (C / 8);
  const int grid = (int)((n_vec + 255) / 256 < 148 * 16 ? (n_vec + 255) / 256 : 148 * 16);
  scale_channels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((__half*)x, n_vec, HW, C, gate);
  return check_launch("scale_channels_kernel");
}
